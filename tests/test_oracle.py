"""CPU tests: the C restatement (oracle/) against the committed golden vectors that came from the unmodified reference,
and -- where oracle/_ref has been built -- against the live reference."""
import os

import numpy as np
import pytest

from conftest import GOLD, load_golden, pair_set
from oracle import oracle as O
from oracle import refharness as RH
from oxdna_b200 import io as oio
from oxdna_b200.sim import parse_temperature

CASES = ["force_field_dna/ref_dna2_nomesh", "lattice8", "lattice27_dense"]


def _params(g):
    return O.dna2_params(parse_temperature(str(g["T"])), float(g["salt"]))


def test_reference_golden_vector_file():
    """test/DNA/FORCE_FIELD/AVG_SEQ/reference.dat of the reference: per-term energies per nucleotide (6 decimals).
    The file was produced with the meshed DNA2 interaction; the analytic form differs in the 6th digit of HB only."""
    ref = np.loadtxt(os.path.join(GOLD, "force_field_dna", "reference_avg_seq.dat"))
    t = oio.read_topology(os.path.join(GOLD, "force_field_dna", "init.top"))
    c = oio.read_conf(os.path.join(GOLD, "force_field_dna", "init.dat"))
    P = O.dna2_params(O.celsius(20.0), 1.0)
    ax = O.axes_from_a1a3(c["a1"], c["a3"])
    pairs = O.verlet_pairs(c["pos"], t["n3"], t["n5"], c["box"], P.rcut + 0.1)
    out = O.forces(P, c["pos"], ax, t["btype"], t["n3"], t["n5"], c["box"], pairs)
    got = out["eterms"] / t["N"]
    assert np.allclose(got, ref, atol=2.5e-6), (got, ref)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_fixture(case):
    g = load_golden(case)
    P = _params(g)
    assert P.rcut == float(g["rcut"])  # bit-equal cutoff: it enters the exact neighbour predicate
    ax = O.axes_from_a1a3(g["a1"], g["a3"])
    pairs = O.verlet_pairs(g["pos"], g["n3"], g["n5"], g["box"], P.rcut + 2 * 0.05)
    assert pair_set(pairs) == pair_set(g["pairs"])
    out = O.forces(P, g["pos"], ax, g["btype"], g["n3"], g["n5"], g["box"], pairs)
    assert np.abs(out["eterms"] - g["energy_split"]).max() < 1e-10
    assert abs(out["U"] - float(g["U"])) < 1e-9
    assert np.abs(out["force"] - g["force"]).max() < 1e-9
    assert np.abs(out["torque_lab"] - g["torque_lab"]).max() < 1e-9
    assert np.abs(out["torque_body"] - g["torque_body"]).max() < 1e-9


@pytest.mark.parametrize("case", ["lattice8", "lattice27_dense"])
def test_oracle_nve_matches_reference_fixture(case):
    g = load_golden(case)
    P = _params(g)
    md = O.MD(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), g["vel"], g["L"], g["btype"], g["n3"], g["n5"], g["box"], 0.003, 0.05)
    md.step(int(g["nve_steps"]))
    assert np.abs(md.pos - g["pos1"]).max() < 1e-9
    assert np.abs(md.vel - g["vel1"]).max() < 1e-9
    assert np.abs(md.L - g["L1"]).max() < 1e-9
    assert np.abs(md.axes[:, 0:3] - g["a11"]).max() < 1e-9


def test_oracle_external_forces_fixture():
    g = load_golden("lattice8_ext")
    P = _params(g)
    ext = [dict(type="mutual_trap", particle=0, ref_particle=39, stiff=0.1, r0=1.2, PBC=1),
           dict(type="mutual_trap", particle=39, ref_particle=0, stiff=0.1, r0=1.2, PBC=1),
           dict(type="trap", particle=45, pos0=(5.0, 5.0, 5.0), stiff=0.5, rate=0.001, dir=(1.0, 0.0, 0.0)),
           dict(type="string", particle=80, F0=0.2, rate=0.0001, dir=(0.0, 1.0, 1.0))]
    md = O.MD(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), g["vel"], g["L"], g["btype"], g["n3"], g["n5"], g["box"], 0.003, 0.05, ext=ext)
    assert np.abs(md.force - g["force"]).max() < 1e-9
    md.step(int(g["nve_steps"]))
    assert np.abs(md.pos - g["pos1"]).max() < 1e-9
    assert np.abs(md.vel - g["vel1"]).max() < 1e-9


def test_thermostat_parameter_restatement():
    T = parse_temperature("300K")
    from oxdna_b200.sim import brownian_params, langevin_params
    assert np.allclose(O.brownian_params(T, 0.003, 103, 0.0, 2.5), brownian_params(T, 0.003, 103, 0.0, 2.5), rtol=1e-14)
    assert np.allclose(O.langevin_params(T, 0.003, 0.0, 2.5), langevin_params(T, 0.003, 0.0, 2.5), rtol=1e-14)
    pt, pr, resc = O.brownian_params(T, 0.003, 103, 0.0, 2.5)
    assert 0 < pr < pt < 1 and abs(resc - np.sqrt(T)) < 1e-15


@pytest.mark.skipif(not RH.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_after_perturbation(tmp_path):
    """Random rigid perturbation of the golden 16-mer, evaluated by both the live reference and the restatement."""
    rng = np.random.default_rng(0)
    top = os.path.join(GOLD, "force_field_dna", "init.top")
    conf = os.path.join(GOLD, "force_field_dna", "init.dat")
    r = RH.Reference(top, conf, interaction_type="DNA2_nomesh", salt_concentration=0.3, T="37C")
    try:
        st = r.state()
        topo = r.topology()
        P = O.dna2_params(O.celsius(37.0), 0.3)
        assert P.rcut == r.rcut()
        for _ in range(3):
            pos = st["pos"] + rng.normal(scale=0.01, size=st["pos"].shape)
            a1 = st["a1"] + rng.normal(scale=0.02, size=st["a1"].shape)
            a3 = st["a3"] + rng.normal(scale=0.02, size=st["a3"].shape)
            ax = O.axes_from_a1a3(a1, a3)
            r.set_state(pos, ax[:, 0:3], ax[:, 6:9])
            ref = r.compute_forces()
            pairs = O.verlet_pairs(pos, topo["n3"], topo["n5"], r.box(), P.rcut + 0.1)
            assert pair_set(pairs) == pair_set(r.pairs())
            out = O.forces(P, pos, ax, topo["btype"], topo["n3"], topo["n5"], r.box(), pairs)
            assert np.abs(out["force"] - ref["force"]).max() < 1e-8 * max(1.0, np.abs(ref["force"]).max())
            assert np.abs(out["torque_lab"] - ref["torque_lab"]).max() < 1e-8 * max(1.0, np.abs(ref["torque_lab"]).max())
            assert abs(out["U"] - ref["U"]) < 1e-9 * max(1.0, abs(ref["U"]))
    finally:
        r.close()
