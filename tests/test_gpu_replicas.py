"""GPU tests of replica batching (oxb_set_replicas: R temperature replicas of one system in ONE context, one launch per kernel for all of
them) and of the batched replica-exchange driver.  Reference behaviour: examples/OXPY_REMD/remd.py runs one process + one context per
replica; a batch must give every replica exactly what its own context at its own temperature gives.

Tolerances: forces / torques 1e-5 * max|.| and energies 1e-6 relative against the ORACLE at each replica's temperature; pair sets of the
hottest replica bit-exact (the batch's lists are built for the hottest Hamiltonian of the ladder: colder replicas get a superset)."""
import numpy as np
import pytest

from conftest import load_golden, pair_set
from oracle import oracle as O
from oxdna_b200 import lattice
from oxdna_b200.remd import ReplicaExchange, geometric_ladder
from oxdna_b200.sim import ReplicaBatch, Simulation, make_batches, parse_temperature

pytestmark = pytest.mark.gpu

TEMPS_K = [290.0, 310.0, 350.0]


def _inp(**over):
    inp = dict(backend="CUDA", interaction_type="DNA2", salt_concentration=0.5, dt=0.003, verlet_skin=0.05, thermostat="no",
               CUDA_sort_every=1, use_edge=1, seed=11)
    inp.update(over)
    return inp


def _replica_confs(g, R, seed=3):
    """R different configurations of the fixture's system: replica r is the fixture rigidly shifted and with its own momenta"""
    rng = np.random.default_rng(seed)
    confs = []
    for r in range(R):
        shift = rng.uniform(-3, 3, size=3) * (r > 0)
        confs.append(dict(box=g["box"], pos=g["pos"] + shift, a1=g["a1"], a3=g["a3"], vel=g["vel"] * (1 + 0.1 * r), L=g["L"] * (1 - 0.1 * r)))
    return confs


def _oracle(g, conf, T):
    P = O.dna2_params(T, float(g["salt"]))
    pairs = O.verlet_pairs(conf["pos"], g["n3"], g["n5"], g["box"], P.rcut + 0.1)
    return O.forces(P, conf["pos"], O.axes_from_a1a3(conf["a1"], conf["a3"]), g["btype"], g["n3"], g["n5"], g["box"], pairs), pairs


@pytest.mark.parametrize("case", ["lattice8", "lattice27_dense"])
@pytest.mark.parametrize("use_edge", [0, 1])
@pytest.mark.parametrize("sort_every", [0, 1])
def test_batch_forces_and_energies_match_the_oracle_per_replica(case, use_edge, sort_every):
    g = load_golden(case)
    topo = dict(btype=g["btype"], n3=g["n3"], n5=g["n5"], strand=g["strand"])
    temps = [parse_temperature(f"{t}K") for t in TEMPS_K]
    confs = _replica_confs(g, len(temps))
    n = len(g["btype"])
    batch = ReplicaBatch(_inp(use_edge=use_edge, CUDA_sort_every=sort_every, salt_concentration=float(g["salt"])), topo, confs, temps)
    try:
        out = batch.ctx.get_forces()
        U = batch.energies()
        pairs = batch.ctx.get_pairs()
        assert np.all(pairs[:, 0] // n == pairs[:, 1] // n), "a Verlet pair crosses replicas"
        for r, T in enumerate(temps):
            ref, ref_pairs = _oracle(g, confs[r], T)
            sl = slice(r * n, (r + 1) * n)
            fmax, tmax = np.linalg.norm(ref["force"], axis=1).max(), np.linalg.norm(ref["torque_lab"], axis=1).max()
            assert np.linalg.norm(out["force"][sl] - ref["force"], axis=1).max() <= 1e-5 * fmax, r
            assert np.linalg.norm(out["torque_lab"][sl] - ref["torque_lab"], axis=1).max() <= 1e-5 * tmax, r
            assert abs(U[r] - ref["U"]) <= 1e-6 * abs(ref["U"]), (r, U[r], ref["U"])
            mine = pair_set(pairs[pairs[:, 0] // n == r] - r * n)
            if r == len(temps) - 1:
                assert mine == pair_set(ref_pairs)  # hottest replica: its own radius, bit-exact
            else:
                assert pair_set(ref_pairs) <= mine
        # the partner-Hamiltonian evaluation of the exchange rule: every replica at its lower neighbour's temperature
        alt = [temps[0]] + temps[:-1]
        U_alt = batch.energies_at(alt)
        for r, T in enumerate(alt):
            ref, _ = _oracle(g, confs[r], T)
            assert abs(U_alt[r] - ref["U"]) <= 1e-6 * abs(ref["U"]), r
        batch.set_temperatures(temps)
        assert np.allclose(batch.energies(), U, rtol=1e-7)
    finally:
        batch.close()


@pytest.mark.parametrize("use_edge", [0, 1])
def test_batch_nve_trajectories_match_single_contexts(use_edge):
    """300 NVE steps with re-sorts and list rebuilds: every replica of the batch follows the trajectory of its own single context
    (FP32 force round-off differs only through the summation order of the edge-centric atomics)."""
    g = load_golden("lattice8")
    topo = dict(btype=g["btype"], n3=g["n3"], n5=g["n5"], strand=g["strand"])
    temps = [parse_temperature(f"{t}K") for t in TEMPS_K]
    confs = _replica_confs(g, len(temps))
    batch = ReplicaBatch(_inp(use_edge=use_edge), topo, confs, temps)
    try:
        batch.run(300)
        states = batch.get_states()
        assert batch.ctx.stats()["n_list_updates"] > 3
    finally:
        batch.close()
    for r, T in enumerate(temps):
        sim = Simulation(_inp(use_edge=use_edge, T=T), topo, confs[r])
        try:
            sim.run(300)
            st = sim.ctx.get_state()
        finally:
            sim.close()
        for k, tol in (("pos", 2e-5), ("a1", 2e-5), ("vel", 2e-4), ("L", 2e-4)):
            assert np.abs(st[k] - states[r][k]).max() < tol, (r, k, np.abs(st[k] - states[r][k]).max())


def test_batch_thermostat_is_per_replica():
    """Brownian thermostat with per-replica constants: <K/N> = 3 T_r for every replica of one batch (equipartition, as the reference's
    THERMOSTATS tests check for one system)."""
    sysm = lattice.duplex_lattice(64, bp=20, spacing=10.0, seed=9)
    temps = [parse_temperature("280K"), parse_temperature("360K")]
    n = len(sysm["pos"])
    confs = []
    for r, T in enumerate(temps):
        v, L = lattice.maxwell_velocities(n, 0.5 * T, 3 + r)
        confs.append(dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L))
    batch = ReplicaBatch(_inp(thermostat="brownian", newtonian_steps=20, pt=0.2, seed=5), sysm, confs, temps)
    try:
        batch.run(8000)
        ks = []
        for _ in range(40):
            batch.run(250)
            ks.append([0.5 * (np.sum(s["vel"] ** 2) + np.sum(s["L"] ** 2)) / n for s in batch.get_states()])
        k = np.mean(ks, axis=0)
        for r, T in enumerate(temps):
            assert abs(k[r] - 3 * T) < 0.03 * 3 * T, (r, k[r], 3 * T)
    finally:
        batch.close()


def test_batched_exchange_equals_one_context_per_replica():
    """The replica-exchange driver over ONE batch takes the same decisions as over one context per replica (the reference's layout):
    NVE, so both follow the same trajectories; 4 replicas, 3 rounds of 60 steps; energies agree to FP32 summation order."""
    g = load_golden("lattice8")
    topo = dict(btype=g["btype"], n3=g["n3"], n5=g["n5"], strand=g["strand"])
    ladder = geometric_ladder(300.0, 306.0, 4) * 0.1 / 300.0
    confs = _replica_confs(g, 4)
    batches = make_batches(_inp(), topo, confs, ladder, ladder_max=float(ladder[-1]))
    singles = [Simulation(_inp(T=float(ladder[r])), topo, confs[r]) for r in range(4)]
    try:
        assert len(batches) == 1
        a = ReplicaExchange(batches, ladder, seed=4)
        b = ReplicaExchange(singles, ladder, seed=4, concurrent=False)
        for _ in range(3):
            a.advance(60)
            b.advance(60)
            acc_a, acc_b = a.exchange(), b.exchange()
            assert acc_a == acc_b
        assert a.location.tolist() == b.location.tolist()
        Ua = batches[0].energies()
        Ub = np.array([s.system_energy() for s in singles])
        assert np.allclose(Ua, Ub, rtol=2e-5)
    finally:
        for s in singles:
            s.close()
        for s in batches:
            s.close()


def test_two_batches_on_one_gpu():
    """make_batches splits replicas that do not fit one 22-bit particle index space; the driver advances the batches from two host threads"""
    g = load_golden("lattice8")
    topo = dict(btype=g["btype"], n3=g["n3"], n5=g["n5"], strand=g["strand"])
    ladder = geometric_ladder(300.0, 330.0, 4) * 0.1 / 300.0
    confs = _replica_confs(g, 4)
    n = len(g["btype"])
    batches = make_batches(_inp(thermostat="brownian", newtonian_steps=53, diff_coeff=2.5), topo, confs, ladder, ladder_max=float(ladder[-1]), max_particles=2 * n)
    try:
        assert [b.n_replicas for b in batches] == [2, 2]
        rx = ReplicaExchange(batches, ladder, seed=1)
        rx.run(3, 100)
        assert sorted(rx.location.tolist()) == [0, 1, 2, 3]
        temps_now = np.concatenate([b.temps for b in batches])
        assert np.allclose(temps_now, ladder[rx.location])
        assert all(np.isfinite(b.energies()).all() for b in batches)
    finally:
        for b in batches:
            b.close()
