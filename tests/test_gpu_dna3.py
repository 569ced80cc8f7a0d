"""GPU parity of oxDNA3 (interaction_type = DNA3) through the C ABI: oxb_set_model_dna3 + the particle-centric kernel of
csrc/forces_dna3.cu against fixtures written by the unmodified reference CPU class DNA3Interaction_nomesh (the class the reference's
CUDA backend instantiates, src/Interactions/InteractionFactory.cpp:63-65) and against the oracle (oracle/oxdna3_oracle.inc)."""
import numpy as np
import pytest

from conftest import load_golden, pair_set
from oracle import oracle as O
from oxdna_b200.sim import Simulation

pytestmark = pytest.mark.gpu
CASES = ["dna3_lattice8", "dna3_lattice27_dense"]


def make_sim(g, topo=None, **over):
    inp = dict(backend="CUDA", interaction_type="DNA3", T=str(g["T"]), salt_concentration=float(g["salt"]), dt=0.003, verlet_skin=0.05,
               thermostat="no", CUDA_sort_every=0, use_edge=0, seed=11, dna3_tables=g["dna3_tables"], dna3_scalars=g["dna3_scalars"])
    inp.update(over)
    topo = topo or dict(btype=g["btype"], n3=g["n3"], n5=g["n5"], strand=g["strand"])
    conf = dict(box=g["box"], pos=g["pos"], a1=g["a1"], a3=g["a3"], vel=g["vel"], L=g["L"])
    return Simulation(inp, topo, conf)


def check_forces(out, ref, tol=1e-5):
    fmax = np.linalg.norm(ref["force"], axis=1).max()
    tmax = np.linalg.norm(ref["torque_lab"], axis=1).max()
    assert np.linalg.norm(out["force"] - ref["force"], axis=1).max() <= tol * fmax
    assert np.linalg.norm(out["torque_lab"] - ref["torque_lab"], axis=1).max() <= tol * tmax
    assert np.linalg.norm(out["torque_body"] - ref["torque_body"], axis=1).max() <= tol * tmax
    assert abs(out["U"] - float(ref["U"])) <= 1e-6 * abs(float(ref["U"]))


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("use_edge", [0, 1])
@pytest.mark.parametrize("sort_every", [0, 1])
@pytest.mark.parametrize("precision", ["mixed", "float"])
def test_dna3_forces_torques_energy_vs_reference(case, use_edge, sort_every, precision):
    """forces, lab and body torques <= 1e-5 max|.| (float: 1e-4), U <= 1e-6, per-term energies, HB energy, pair set bit-exact"""
    g = load_golden(case)
    sim = make_sim(g, use_edge=use_edge, CUDA_sort_every=sort_every, backend_precision=precision)
    try:
        assert pair_set(sim.ctx.get_pairs()) == pair_set(g["pairs"])
        out = sim.ctx.get_forces()
        check_forces(out, g, tol=1e-5 if precision == "mixed" else 1e-4)
        U, K = sim.ctx.energy()
        assert abs(U - float(g["U"])) <= 1e-6 * abs(float(g["U"]))
        split = sim.ctx.energy_split()
        assert np.abs(np.asarray(split)[:8] - g["energy_split"]).max() <= 2e-6 * abs(float(g["U"]))
        hb = out["hb_energy"].sum() * 0.5
        assert abs(hb - float(g["energy_split"][4])) <= 1e-5 * abs(float(g["energy_split"][4])) + 1e-6
    finally:
        sim.close()


@pytest.mark.parametrize("case,sort_every", [("dna3_lattice8", 0), ("dna3_lattice27_dense", 1)])
def test_dna3_nve_trajectory_vs_reference(case, sort_every):
    """100 NVE steps through oxb_run (CUDA-graph batches, device-side list staleness) against the reference CPU run"""
    g = load_golden(case)
    sim = make_sim(g, CUDA_sort_every=sort_every)
    try:
        sim.run(int(g["nve_steps"]))
        st = sim.ctx.get_state()
        assert np.abs(st["pos"] - g["pos1"]).max() < 2e-4
        assert np.abs(st["vel"] - g["vel1"]).max() < 2e-3
        assert np.abs(st["L"] - g["L1"]).max() < 2e-3
        U, K = sim.ctx.energy()
        assert abs(U - float(g["U1"])) <= 2e-5 * abs(float(g["U1"]))
    finally:
        sim.close()


def test_dna3_nicked_strands_coaxial_stacking_vs_oracle():
    """strands nicked in the middle and next to an end: the three K branches of the coaxial-stacking term (DNA3Interaction.cpp:1814-1825),
    strand-end types (tetramer index 5) in stacking, FENE and cross stacking"""
    g = load_golden("dna3_lattice27_dense")
    n3, n5 = g["n3"].copy(), g["n5"].copy()
    starts = np.flatnonzero(g["n3"] < 0)
    for k, first in enumerate(starts[::3]):
        i = first + (1 if k % 3 == 0 else 9)
        j = n5[i]
        n5[i], n3[j] = -1, -1
    sim = make_sim(g, topo=dict(btype=g["btype"], n3=n3, n5=n5, strand=g["strand"]), CUDA_sort_every=1)
    try:
        pairs = sim.ctx.get_pairs()
        P = O.dna3_params(g["dna3_tables"], g["dna3_scalars"])
        ax = O.axes_from_a1a3(g["a1"], g["a3"])
        assert pair_set(pairs) == pair_set(O.verlet_pairs(g["pos"], n3, n5, g["box"], P.rcut + 0.1))
        ref = O.forces(P, g["pos"], ax, g["btype"], n3, n5, g["box"], pairs)
        assert abs(ref["eterms"][6]) > 1e-2
        check_forces(sim.ctx.get_forces(), ref)
        assert np.abs(np.asarray(sim.ctx.energy_split())[:8] - ref["eterms"]).max() <= 2e-6 * abs(ref["U"])
    finally:
        sim.close()


def test_dna3_thermostatted_run_is_deterministic():
    """Brownian thermostat + Hilbert re-sorts: two runs of 300 steps agree bit for bit (the particle-centric kernel has no atomics); three runs
    of 100 (other batch boundaries: the step at a boundary goes through the single-phase integrator instantiations, and lists may be rebuilt
    at other steps, which reorders the FP32 sums) agree to round-off; the potential energy stays in the thermal band of the fixture"""
    g = load_golden("dna3_lattice8")
    over = dict(thermostat="brownian", newtonian_steps=103, diff_coeff=2.5, CUDA_sort_every=1)
    a, b, c = make_sim(g, **over), make_sim(g, **over), make_sim(g, **over)
    try:
        a.run(300)
        b.run(300)
        for _ in range(3):
            c.run(100)
        sa, sb, sc = a.ctx.get_state(), b.ctx.get_state(), c.ctx.get_state()
        assert np.array_equal(sa["pos"], sb["pos"]) and np.array_equal(sa["vel"], sb["vel"]) and np.array_equal(sa["L"], sb["L"])
        assert np.abs(sa["pos"] - sc["pos"]).max() < 1e-5 and np.abs(sa["vel"] - sc["vel"]).max() < 1e-4
        U, K = a.ctx.energy()
        assert abs(U / len(g["pos"]) - float(g["U"]) / len(g["pos"])) < 0.08
    finally:
        a.close()
        b.close()
        c.close()


def test_dna3_model_switch_on_one_context():
    """a context that ran oxDNA2 with the edge pipeline takes the oxDNA3 model (particle-centric pass) and goes back"""
    g = load_golden("dna3_lattice8")
    sim = make_sim(g, interaction_type="DNA2", use_edge=1)
    try:
        U2 = sim.ctx.energy()[0]
        sim.ctx.set_model_dna3(g["dna3_tables"], g["dna3_scalars"])
        out = sim.ctx.get_forces()
        check_forces(out, g)
        sim._set_model()
        assert abs(sim.ctx.energy()[0] - U2) <= 1e-6 * abs(U2)
    finally:
        sim.close()


def test_dna3_full_size_c2_geometry_against_the_oracle():
    """81,920 nucleotides (the C2 lattice, L = 130) under oxDNA3 with the sequence-dependent tables of the fixture, after 300 thermalising
    steps with Hilbert re-sorts: pair set bit-exact (rcut of oxDNA3), forces, lab torques and energy against the oracle on the downloaded state"""
    from oxdna_b200 import lattice
    from oxdna_b200.sim import parse_temperature
    g = load_golden("dna3_lattice8")
    sysm = lattice.duplex_lattice(2048, bp=20, spacing=10.0, seed=12345)
    N = len(sysm["pos"])
    v, L = lattice.maxwell_velocities(N, parse_temperature("300K"), 5)
    inp = dict(backend="CUDA", interaction_type="DNA3", T="300K", salt_concentration=0.5, dt=0.003, verlet_skin=0.05, thermostat="brownian",
               newtonian_steps=103, diff_coeff=2.5, CUDA_sort_every=1, use_edge=1, seed=11, dna3_tables=g["dna3_tables"], dna3_scalars=g["dna3_scalars"])
    sim = Simulation(inp, sysm, dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L))
    try:
        sim.run(300)
        st = sim.ctx.get_state()
        P = O.dna3_params(g["dna3_tables"], g["dna3_scalars"])
        pairs = O.verlet_pairs(st["pos"], sysm["n3"], sysm["n5"], sysm["box"], P.rcut + 2 * 0.05)
        sim.ctx.update_lists()
        assert pair_set(sim.ctx.get_pairs()) == pair_set(pairs)
        ref = O.forces(P, st["pos"], O.axes_from_a1a3(st["a1"], st["a3"]), sysm["btype"], sysm["n3"], sysm["n5"], sysm["box"], pairs)
        sim.ctx.compute_forces()
        check_forces(sim.ctx.get_forces(), ref)
        assert np.abs(np.asarray(sim.ctx.energy_split())[:8] - ref["eterms"]).max() <= 2e-6 * abs(ref["U"])
    finally:
        sim.close()


def test_dna3_refuses_what_it_does_not_serve():
    """replica batches and a temperature change without new tables are explicit errors (the tables depend on T and are input)"""
    from oxdna_b200 import capi
    g = load_golden("dna3_lattice8")
    sim = make_sim(g)
    try:
        with pytest.raises(ValueError):
            sim.update_temperature("310K")
        sim.update_temperature("300K", g["dna3_tables"], g["dna3_scalars"])
        check_forces(sim.ctx.get_forces(), g)
    finally:
        sim.close()
    c = capi.Context(len(g["pos"]))
    try:
        c.set_replicas(2)
        with pytest.raises(capi.OxbError):
            c.set_model_dna3(g["dna3_tables"], g["dna3_scalars"])
    finally:
        c.close()


@pytest.mark.parametrize("use_edge", [0, 1])
def test_dna3_special_base_types_vs_oracle(use_edge):
    """dummy bases (btype = type = 4) and a custom pair -300 / 303 with hb_multiplier, both force variants, against the oracle"""
    from conftest import dna3_special_types
    g = load_golden("dna3_lattice8")
    bt, sc = dna3_special_types(g)
    sim = make_sim(g, topo=dict(btype=bt, n3=g["n3"], n5=g["n5"], strand=g["strand"]), dna3_scalars=sc, use_edge=use_edge, CUDA_sort_every=1)
    try:
        P = O.dna3_params(g["dna3_tables"], sc)
        ref = O.forces(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), bt, g["n3"], g["n5"], g["box"], g["pairs"])
        assert abs(ref["eterms"][4] - float(g["energy_split"][4])) > 0.5
        check_forces(sim.ctx.get_forces(), ref)
        assert np.abs(np.asarray(sim.ctx.energy_split())[:8] - ref["eterms"]).max() <= 2e-6 * abs(ref["U"])
    finally:
        sim.close()


@pytest.mark.parametrize("use_edge", [0, 1])
def test_dna3_average_sequence_tables_vs_oracle(use_edge):
    """the average-sequence tables bench.py --workload c2_dna3 / c4_dna3 runs on (pinned to the live reference in test_oracle_dna3.py)"""
    import os
    g = load_golden("dna3_lattice8")
    avg = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dna3_tables_avg_300K_salt05.npz"))
    sim = make_sim(g, dna3_tables=avg["dna3_tables"], dna3_scalars=avg["dna3_scalars"], use_edge=use_edge, CUDA_sort_every=1)
    try:
        P = O.dna3_params(avg["dna3_tables"], avg["dna3_scalars"])
        pairs = O.verlet_pairs(g["pos"], g["n3"], g["n5"], g["box"], P.rcut + 0.1)
        assert pair_set(sim.ctx.get_pairs()) == pair_set(pairs)
        ref = O.forces(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), g["btype"], g["n3"], g["n5"], g["box"], pairs)
        check_forces(sim.ctx.get_forces(), ref)
    finally:
        sim.close()
