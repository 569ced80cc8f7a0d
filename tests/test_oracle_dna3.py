"""CPU tests of the oxDNA3 restatement (oracle/oxdna3_oracle.inc) against fixtures written by the unmodified reference CPU class
DNA3Interaction_nomesh (the class the CUDA backend instantiates: src/Interactions/InteractionFactory.cpp:63-65) and -- where
oracle/_ref has been built -- against the live class.  The tetramer-indexed parameter tables are part of the fixtures: they are what
DNA3Interaction::init leaves in the class, i.e. what the reference's CUDA class uploads (CUDADNA3Interaction.cu:46-150)."""
import os

import numpy as np
import pytest

from conftest import load_golden, pair_set
from oracle import oracle as O
from oracle import refharness as RH
from oxdna_b200 import io as oio

CASES = ["dna3_lattice8", "dna3_lattice27_dense"]
SEQ = "/root/reference/oxDNA3_sequence_dependent_parameters.txt"


def _params(g):
    return O.dna3_params(g["dna3_tables"], g["dna3_scalars"])


@pytest.mark.parametrize("case", CASES)
def test_dna3_oracle_matches_reference_fixture(case):
    g = load_golden(case)
    P = _params(g)
    assert P.rcut == float(g["rcut"])
    ax = O.axes_from_a1a3(g["a1"], g["a3"])
    pairs = O.verlet_pairs(g["pos"], g["n3"], g["n5"], g["box"], P.rcut + 2 * 0.05)
    assert pair_set(pairs) == pair_set(g["pairs"])
    out = O.forces(P, g["pos"], ax, g["btype"], g["n3"], g["n5"], g["box"], pairs)
    assert np.abs(out["eterms"] - g["energy_split"]).max() < 1e-10
    assert abs(out["U"] - float(g["U"])) < 1e-9
    # every term of the model is exercised (bonded excluded volume and coaxial stacking only on the dense case)
    on = np.abs(g["energy_split"]) > 0
    assert on[[0, 2, 4, 5, 7]].all()
    assert np.abs(out["force"] - g["force"]).max() < 1e-9
    assert np.abs(out["torque_lab"] - g["torque_lab"]).max() < 1e-9
    assert np.abs(out["torque_body"] - g["torque_body"]).max() < 1e-9


@pytest.mark.parametrize("case", CASES)
def test_dna3_oracle_nve_matches_reference_fixture(case):
    g = load_golden(case)
    P = _params(g)
    md = O.MD(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), g["vel"], g["L"], g["btype"], g["n3"], g["n5"], g["box"], 0.003, 0.05)
    md.step(int(g["nve_steps"]))
    assert np.abs(md.pos - g["pos1"]).max() < 1e-9
    assert np.abs(md.vel - g["vel1"]).max() < 1e-9
    assert np.abs(md.L - g["L1"]).max() < 1e-9
    assert np.abs(md.axes[:, 0:3] - g["a11"]).max() < 1e-9


def test_dna3_reference_force_is_not_the_gradient_in_the_stacking_phi_terms():
    """The reference (CPU class and CUDA kernel alike) differentiates cos(phi1), cos(phi2) of the stacking term with the oxDNA2 lever
    gamma = POS_STACK - POS_BACK = 0.74 (DNA3Interaction.cpp:1363, CUDA_DNA3.cuh:510) while the oxDNA3 stacking site sits at 0.37:
    ref_form = 1 restates that literally (and matches the fixtures), ref_form = 0 is the gradient of the same energy."""
    g = load_golden("dna3_lattice8")
    P = _params(g)
    ax = O.axes_from_a1a3(g["a1"], g["a3"])
    args = (ax, g["btype"], g["n3"], g["n5"], g["box"], g["pairs"])
    ref = O.forces(P, g["pos"], *args)
    P.ref_form = 0
    grad = O.forces(P, g["pos"], *args)
    assert np.abs(grad["eterms"] - ref["eterms"]).max() == 0.0
    d = np.abs(grad["force"] - ref["force"]).max()
    assert 1e-4 < d < 0.1
    # central differences of the energy reproduce the gradient form, not the reference's
    rng = np.random.default_rng(0)
    for i in rng.choice(len(g["pos"]), 6, replace=False):
        for k in range(3):
            h = 1e-6
            pp, pm = g["pos"].copy(), g["pos"].copy()
            pp[i, k] += h
            pm[i, k] -= h
            fd = -(O.forces(P, pp, *args)["U"] - O.forces(P, pm, *args)["U"]) / (2 * h)
            assert abs(fd - grad["force"][i, k]) < 2e-7


@pytest.mark.skipif(not (RH.available() and os.path.exists(SEQ)), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("T,salt,mbf", [("300K", 0.5, None), ("37C", 0.15, 10.0)])
def test_dna3_oracle_matches_live_reference_after_perturbation(tmp_path, T, salt, mbf):
    """Strong random rigid perturbation of the dense fixture (excluded volume, coaxial stacking and stretched bonds in range), evaluated
    by the live DNA3Interaction_nomesh and by the restatement fed with the tables dumped from that very object.  Pairs inside the coaxial
    window agree to the accuracy of the reference's theta4/5/6 MESHES (DNA3Interaction_nomesh inherits the meshed scalar f4 of
    DNA2Interaction for coaxial stacking, DNA3Interaction.cpp:1805-1808); everything else to 1e-11."""
    g = load_golden("dna3_lattice27_dense")
    rng = np.random.default_rng(11)
    N = len(g["pos"])
    amp = 0.3 if mbf is None else 1.0  # without max_backbone_force the reference refuses bonds outside the FENE range
    pos = g["pos"] + rng.normal(0, 0.05 * amp, (N, 3))
    a1 = g["a1"] + rng.normal(0, 0.1 * amp, (N, 3))
    a3 = g["a3"] + rng.normal(0, 0.1 * amp, (N, 3))
    ax = O.axes_from_a1a3(a1, a3)
    # nick every third strand in the middle: the two stacked neighbours become a non-bonded pair inside the coaxial-stacking window,
    # with and without flanking neighbours on the far side (the three K branches of DNA3Interaction.cpp:1814-1825)
    n3, n5 = g["n3"].copy(), g["n5"].copy()
    starts = np.flatnonzero(g["n3"] < 0)
    for s_i, first in enumerate(starts[::3]):
        i = first + (1 if s_i % 3 == 0 else 9)   # 3' end particle has n3 = -1; walk 5'-wards: i's n5 is i + 1 in the lattice generator
        j = n5[i]
        n5[i], n3[j] = -1, -1
    top, conf = str(tmp_path / "t.top"), str(tmp_path / "t.dat")
    oio.write_topology(top, g["btype"], n3, n5, g["strand"])
    oio.write_conf(conf, g["box"], pos, ax[:, 0:3], ax[:, 6:9], g["vel"], g["L"])
    keys = dict(interaction_type="DNA3_nomesh", salt_concentration=salt, T=T, use_average_seq=0, seq_dep_file=SEQ)
    if mbf is not None:
        keys.update(max_backbone_force=mbf)
    r = RH.Reference(top, conf, **keys)
    try:
        tab, sc = np.zeros((215, 900)), np.zeros(40)
        k = RH.lib().oxref_dna3_tables(RH._p(tab), RH._p(sc))
        assert k == O.DNA3_NSCALARS
        st = r.state()
        ref = r.compute_forces()
        split = r.energy_split()
        pairs = r.pairs()
        topo = r.topology()
        box = r.box()
    finally:
        r.close()
    P = O.dna3_params(tab, sc[:k])
    ax = O.axes_from_a1a3(st["a1"], st["a3"])
    out = O.forces(P, st["pos"], ax, topo["btype"], topo["n3"], topo["n5"], box, pairs)
    d = np.abs(out["eterms"] - split)
    assert abs(split[6]) > 1e-3, "coaxial stacking must be active"
    assert np.delete(d, 6).max() < 1e-9, d
    assert d[6] < 2e-4 * abs(split[6]), d          # meshed theta4/5/6 in the CPU class, analytic in the restatement (and on the device)
    fm = np.abs(ref["force"]).max()
    assert np.abs(out["force"] - ref["force"]).max() < 2e-4 * fm
    # particles of pairs without coaxial stacking agree to rounding: find them through the restatement itself
    cx = _coaxial_particles(P, st, ax, topo, box, pairs)
    quiet = np.setdiff1d(np.arange(N), cx)
    assert len(quiet) > N // 2
    assert np.abs(out["force"][quiet] - ref["force"][quiet]).max() < 1e-9
    assert np.abs(out["torque_body"][quiet] - ref["torque_body"][quiet]).max() < 1e-9
    # with the reference's 6-interval cubic meshes for those three factors (cxst_mesh = 1) the restatement IS the CPU class, coaxial term included
    P.cxst_mesh = 1
    m = O.forces(P, st["pos"], ax, topo["btype"], topo["n3"], topo["n5"], box, pairs)
    assert np.abs(m["eterms"] - split).max() < 1e-9
    assert np.abs(m["force"] - ref["force"]).max() < 1e-9
    assert np.abs(m["torque_body"] - ref["torque_body"]).max() < 1e-9


def _coaxial_particles(P, st, ax, topo, box, pairs):
    """particles that belong to a pair with a non-zero coaxial-stacking energy (evaluated pair by pair through the oracle)"""
    hit = set()
    for a, b in pairs:
        e = O.forces(P, st["pos"], ax, topo["btype"], topo["n3"], topo["n5"], box, np.array([[a, b]], dtype=np.int32))["eterms"]
        if e[6] != 0.0:
            hit.update((int(a), int(b)))
    return np.array(sorted(hit), dtype=int)


@pytest.mark.skipif(not (RH.available() and os.path.exists(SEQ)), reason="oracle/_ref not built (needs /root/reference)")
def test_dna3_special_base_types_against_the_live_reference(tmp_path):
    """Base types outside 0..3: a dummy base (btype 4: stacking / base sites at the oxDNA2 offsets 0.34 / 0.40, DNANucleotide.cpp:71-79) and a
    custom pair 303 / -300 (hydrogen-bonds as T-A, |btype| >= 300: strength times hb_multiplier, DNA3Interaction.cpp:1483-1484): the oracle
    against the live DNA3Interaction_nomesh (the device formulation is held to the oracle on the same construction: test_host_model.py, test_gpu_dna3.py)"""
    g = load_golden("dna3_lattice8")
    bt = g["btype"].copy()
    N = len(bt)
    # duplex 0: strand 0 = particles 0..19, strand 1 = 20..39, base i pairs with 39 - i
    n_special = 0
    for i in range(20):
        j = 39 - i
        if (bt[i], bt[j]) == (3, 0):
            bt[i], bt[j] = 303, -300  # type_of(303) = 3 (T), type_of(-300) = 0 (A)
            n_special += 1
        elif (bt[i], bt[j]) == (0, 3):
            bt[i], bt[j] = -300, 303
            n_special += 1
    assert n_special >= 3
    bt[45], bt[130] = 4, 4
    top, conf = str(tmp_path / "t.top"), str(tmp_path / "t.dat")
    oio.write_topology(top, bt, g["n3"], g["n5"], g["strand"])
    oio.write_conf(conf, g["box"], g["pos"], g["a1"], g["a3"], g["vel"], g["L"])
    r = RH.Reference(top, conf, interaction_type="DNA3_nomesh", salt_concentration=0.5, T="300K", use_average_seq=0, seq_dep_file=SEQ, hb_multiplier=1.7)
    try:
        tab, sc = np.zeros((215, 900)), np.zeros(40)
        k = RH.lib().oxref_dna3_tables(RH._p(tab), RH._p(sc))
        st, ref, split, pairs, topo, box = r.state(), r.compute_forces(), r.energy_split(), r.pairs(), r.topology(), r.box()
    finally:
        r.close()
    assert sc[4] == 1.7 and (topo["btype"] == bt).all()
    P = O.dna3_params(tab, sc[:k])
    ax = O.axes_from_a1a3(st["a1"], st["a3"])
    out = O.forces(P, st["pos"], ax, bt, g["n3"], g["n5"], box, pairs)
    assert np.abs(out["eterms"] - split).max() < 1e-9
    assert np.abs(out["force"] - ref["force"]).max() < 1e-9
    assert np.abs(out["torque_body"] - ref["torque_body"]).max() < 1e-9
    # the multiplier and the dummy sites matter: with plain types the hydrogen-bonding energy is smaller
    plain = O.forces(P, st["pos"], ax, g["btype"], g["n3"], g["n5"], box, pairs)
    assert out["eterms"][4] < plain["eterms"][4] - 0.5
    assert np.abs(out["force"] - plain["force"])[[45, 130]].max() > 1e-3  # the dummy bases sit elsewhere and carry other parameters


_AVG_SCRIPT = r"""
import os, sys
import numpy as np
root, tmp = sys.argv[1], sys.argv[2]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from conftest import load_golden
from oracle import oracle as O, refharness as RH
from oxdna_b200 import io as oio
g = load_golden("dna3_lattice8")
avg = np.load(os.path.join(root, "tests", "golden", "dna3_tables_avg_300K_salt05.npz"))
top, conf = os.path.join(tmp, "t.top"), os.path.join(tmp, "t.dat")
oio.write_topology(top, g["btype"], g["n3"], g["n5"], g["strand"])
oio.write_conf(conf, g["box"], g["pos"], g["a1"], g["a3"], g["vel"], g["L"])
r = RH.Reference(top, conf, interaction_type="DNA3_nomesh", salt_concentration=0.5, T="300K")
tab, sc = np.zeros((215, 900)), np.zeros(40)
k = RH.lib().oxref_dna3_tables(RH._p(tab), RH._p(sc))
ref, split, pairs = r.compute_forces(), r.energy_split(), r.pairs()
r.close()
assert np.array_equal(tab, avg["dna3_tables"]) and np.array_equal(sc[:k], avg["dna3_scalars"]), "tables differ"
P = O.dna3_params(avg["dna3_tables"], avg["dna3_scalars"])
out = O.forces(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), g["btype"], g["n3"], g["n5"], g["box"], pairs)
assert np.abs(out["eterms"] - split).max() < 1e-9
assert np.abs(out["force"] - ref["force"]).max() < 1e-9
assert np.abs(out["torque_body"] - ref["torque_body"]).max() < 1e-9
"""


@pytest.mark.skipif(not RH.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_dna3_average_sequence_tables_match_the_live_reference(tmp_path):
    """use_average_seq = 1 (no parameter file): the committed average-sequence tables (tests/golden/dna3_tables_avg_300K_salt05.npz, what
    bench.py --workload c2_dna3 / c4_dna3 feeds the device) are what the live class holds, and the oracle reproduces its forces with them.
    In a FRESH process, as the reference's CLI runs: with use_average_seq = 1 the class never assigns F1_SD_SHIFT and fills the stacking
    F1_SD_EPS in its constructor from a temperature that is not set yet (DNA3Interaction.cpp:82-83), i.e. from whatever the heap holds --
    zeros in a new process (the tables of the fixture, and of the reference CUDA run the bench compares with), leftovers of the previous
    interaction object otherwise."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", _AVG_SCRIPT, root, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:]
