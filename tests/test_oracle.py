"""CPU tests: the C restatement (oracle/) against the committed golden vectors that came from the unmodified reference,
and -- where oracle/_ref has been built -- against the live reference."""
import os

import numpy as np
import pytest

from conftest import GOLD, ext2_forces, ext3_forces, load_golden, pair_set
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


def test_oracle_external_forces_second_batch_fixture():
    """repulsion_plane_moving (range of reference particles, `all` and single), generic_central_force (gravity with both cut-offs, and
    repulsive without), LJ_cone (only_repulsive), com (two forces with their own index lists), yukawa_sphere (WCA branch active),
    repulsive_sphere_moving (growing and translating) -- forces at step 0 and 100 steps against the reference CPU run"""
    g = load_golden("lattice8_ext3")
    P = _params(g)
    ext = ext3_forces(g["pos"])
    ref = g["force"] - g["force_noext"]
    assert (np.abs(ref).max(axis=1) > 1e-9).sum() >= 250
    assert np.abs(O.ext_forces(ext, g["pos"], g["box"], 0) - ref).max() < 1e-9
    # every force type contributes: dropping any entry changes the result
    for k in range(len(ext)):
        assert np.abs(O.ext_forces(ext[:k] + ext[k + 1:], g["pos"], g["box"], 0) - ref).max() > 1e-6, ext[k]["type"]
    md = O.MD(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), g["vel"], g["L"], g["btype"], g["n3"], g["n5"], g["box"], 0.003, 0.05, ext=ext)
    assert np.abs(md.force - g["force"]).max() < 1e-9
    md.step(int(g["nve_steps"]))
    assert np.abs(md.pos - g["pos1"]).max() < 1e-9
    assert np.abs(md.vel - g["vel1"]).max() < 1e-9


def test_oracle_further_external_forces_fixture():
    """repulsion_plane (moving, clamped), attraction_plane (both branches), sphere (shrinking, r_ext), LJ_wall (only_repulsive),
    lowdim_trap (visibility mask), with `particle = all` entries -- forces at step 0 and 100 steps against the reference CPU run"""
    g = load_golden("lattice8_ext2")
    P = _params(g)
    ext = ext2_forces(g["pos"])
    assert np.abs(O.ext_forces(ext, g["pos"], g["box"], 0) - (g["force"] - g["force_noext"])).max() < 1e-9
    md = O.MD(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), g["vel"], g["L"], g["btype"], g["n3"], g["n5"], g["box"], 0.003, 0.05, ext=ext)
    assert np.abs(md.force - g["force"]).max() < 1e-9
    md.step(int(g["nve_steps"]))
    assert np.abs(md.pos - g["pos1"]).max() < 1e-9
    assert np.abs(md.vel - g["vel1"]).max() < 1e-9


def test_oracle_string_force_dir_as_centre():
    """ConstantRateForce with dir_as_centre = true (src/Forces/ConstantRateForce.cpp:54-61; src/CUDA/Backends/CUDA_MD.cuh:114-130): the
    force (F0 + rate * step) points from the particle towards the point given as `dir`."""
    g = load_golden("lattice8")
    ext = [dict(type="string", particle=7, F0=0.3, rate=0.002, dir=(4.0, -2.0, 11.0), dir_as_centre=1),
           dict(type="string", particle=9, F0=0.1, rate=0.0, dir=(0.0, 0.0, 2.0))]
    F = O.ext_forces(ext, g["pos"], g["box"], 50)
    d = np.array([4.0, -2.0, 11.0]) - g["pos"][7]
    assert np.allclose(F[7], (0.3 + 0.002 * 50) * d / np.linalg.norm(d), atol=1e-14)
    assert np.allclose(F[9], [0.0, 0.0, 0.1], atol=1e-15)
    assert np.count_nonzero(np.abs(F).sum(axis=1)) == 2


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


# ---------------------------------------------------------------------------------------------------------------- oxRNA2
def _rna_params(g, cpu_quirks=True):
    mis = float(g["mismatch"]) if "mismatch" in g else -1.0
    P = O.rna2_params(parse_temperature(str(g["T"])), float(g["salt"]), cpu_quirks=cpu_quirks, mismatch_repulsion=mis >= 0,
                      mismatch_repulsion_strength=max(mis, 0.0))
    if "sd_stck" in g:
        O.rna2_params_seqdep(P, g["sd_stck"], float(g["sd_st_t_dep"]), g["sd_cross"], float(g["sd_hb_AT"]), float(g["sd_hb_GC"]), float(g["sd_hb_GT"]))
    return P


def test_rna_reference_golden_vector_file(which="avg_seq"):
    """test/RNA/FORCE_FIELD/AVG_SEQ/reference.dat of the reference: per-term energies per nucleotide (6 decimals).  (The
    reference's SEQ_DEP test sets use_average_seq = true and holds the same numbers; the sequence-dependent tables are
    pinned by the ref_rna2_seqdep fixture below.)  The CPU class meshes the hydrogen-bonding f4 factors; the analytic form
    (what the CUDA kernels evaluate) differs from it in the 6th digit of that term only."""
    d = os.path.join(GOLD, "force_field_rna")
    ref = np.loadtxt(os.path.join(d, f"reference_{which}.dat"))
    g = load_golden("force_field_rna/ref_rna2" + ("_seqdep" if which == "seq_dep" else ""))
    t = oio.read_topology(os.path.join(d, "init.top"))
    c = oio.read_conf(os.path.join(d, "init.dat"))
    for k in ("btype", "n3", "n5"):
        assert (t[k] == g[k]).all()
    P = _rna_params(g)
    ax = O.axes_from_a1a3(c["a1"], c["a3"])
    pairs = O.verlet_pairs(c["pos"], t["n3"], t["n5"], c["box"], P.rcut + 0.1)
    out = O.forces(P, c["pos"], ax, t["btype"], t["n3"], t["n5"], c["box"], pairs)
    assert np.allclose(out["eterms"] / t["N"], ref, atol=2.5e-6), (out["eterms"] / t["N"], ref)


def test_rna_cpu_quirk_term_is_active_on_the_quirks_fixture():
    """tests/golden/rna_quirks.npz (oracle/make_golden.py rna_quirks): a strongly perturbed all-A configuration of the reference's 16-nt
    RNA system on which the reference CPU class's force is visibly NOT the gradient of its energy: the mirrored coaxial theta1 term
    (src/Interactions/RNAInteraction.cpp:1046; cpu_quirks bit 1).  The restatement with cpu_quirks = 3 reproduces the live CPU class
    to rounding; with cpu_quirks = 0 (the gradient, what src/CUDA/Interactions/CUDA_RNA.cuh:896 evaluates) the energy is identical
    and the torques differ by 2.5 %.  The second such spot (phi2 stacking term, RNAInteraction.cpp:620; bit 0) is dormant with the
    stock parameters: inside the theta-B2 window cos(phi2) >= 0.04, where f5 is flat -- it changes nothing here or anywhere we looked."""
    g = load_golden("rna_quirks")
    assert float(g["energy_split"][4]) == 0.0  # no hydrogen bonding: no meshed factor in the CPU numbers
    ax = O.axes_from_a1a3(g["a1"], g["a3"])
    out = {}
    for q in (3, 2, 1, 0):
        P = _rna_params(g, cpu_quirks=q)
        pairs = O.verlet_pairs(g["pos"], g["n3"], g["n5"], g["box"], P.rcut + 2 * 0.05)
        assert pair_set(pairs) == pair_set(g["pairs"])
        out[q] = O.forces(P, g["pos"], ax, g["btype"], g["n3"], g["n5"], g["box"], pairs)
    fmax = np.linalg.norm(g["force"], axis=1).max()
    tmax = np.linalg.norm(g["torque_lab"], axis=1).max()
    assert np.abs(out[3]["force"] - g["force"]).max() < 1e-10 * fmax
    assert np.abs(out[3]["torque_lab"] - g["torque_lab"]).max() < 1e-10 * tmax
    assert abs(out[3]["U"] - float(g["U"])) < 1e-10 * abs(float(g["U"]))
    assert abs(out[0]["U"] - out[3]["U"]) < 1e-12 * abs(out[3]["U"])  # the energy has no quirk
    dT = np.linalg.norm(out[0]["torque_lab"] - out[3]["torque_lab"], axis=1).max() / tmax
    assert dT > 1e-2, dT  # three orders of magnitude above the 1e-5 GPU criterion
    # the phi2 term is dormant: bit 0 changes nothing
    assert np.array_equal(out[2]["torque_lab"], out[3]["torque_lab"]) and np.array_equal(out[1]["torque_lab"], out[0]["torque_lab"])
    assert np.array_equal(out[2]["force"], out[3]["force"])


@pytest.mark.parametrize("case", ["force_field_rna/ref_rna2", "force_field_rna/ref_rna2_seqdep", "rna_lattice8", "rna_lattice8_nohb", "rna_lattice8_seqdep"])
def test_rna_oracle_matches_reference_fixture(case):
    """Fixtures written by the unmodified reference CPU backend (interaction_type = RNA2).  Everything but the meshed
    hydrogen-bonding term agrees to rounding (cpu_quirks: see oxdna_oracle.h); that term agrees to mesh accuracy."""
    g = load_golden(case)
    P = _rna_params(g)
    assert P.rcut == float(g["rcut"])
    ax = O.axes_from_a1a3(g["a1"], g["a3"])
    pairs = O.verlet_pairs(g["pos"], g["n3"], g["n5"], g["box"], P.rcut + 2 * 0.05)
    assert pair_set(pairs) == pair_set(g["pairs"])
    out = O.forces(P, g["pos"], ax, g["btype"], g["n3"], g["n5"], g["box"], pairs)
    d = out["eterms"] - g["energy_split"]
    hb = 4
    assert np.abs(np.delete(d, hb)).max() < 1e-10
    has_hb = abs(g["energy_split"][hb]) > 0
    # the mesh derivative is the least accurate part of the CPU class; mismatched pairs sit anywhere in the angular windows
    # (also on the coarse ends of the meshes), Watson-Crick pairs near the well centres
    # (6- and 12-point meshes, rna_model.h:1124-1129)
    tol = (5e-3 if float(g.get("mismatch", -1.0)) >= 0 else 1e-4) if has_hb else 1e-9
    assert abs(d[hb]) <= 1e-4 * abs(g["energy_split"][hb])
    for k in ("force", "torque_lab", "torque_body"):
        assert np.abs(out[k] - g[k]).max() < tol * max(1.0, np.abs(g[k]).max()), k
    # with the CPU class's cubic meshes restated for the six hydrogen-bonding factors (cpu_quirks bit 2) the restatement reproduces the
    # fixture in EVERY term to rounding -- the analytic form above is what the CUDA kernels (the reference's and ours) evaluate
    P.cpu_quirks = int(P.cpu_quirks) | 4
    m = O.forces(P, g["pos"], ax, g["btype"], g["n3"], g["n5"], g["box"], pairs)
    assert np.abs(m["eterms"] - g["energy_split"]).max() < 1e-10
    for k in ("force", "torque_lab", "torque_body"):
        assert np.abs(m[k] - g[k]).max() < 1e-9 * max(1.0, np.abs(g[k]).max()), k


def test_rna_oracle_nve_matches_reference_fixture():
    g = load_golden("rna_lattice8_nohb")
    P = _rna_params(g)
    md = O.MD(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), g["vel"], g["L"], g["btype"], g["n3"], g["n5"], g["box"], 0.003, 0.05)
    md.step(int(g["nve_steps"]))
    assert np.abs(md.pos - g["pos1"]).max() < 1e-9
    assert np.abs(md.vel - g["vel1"]).max() < 1e-9
    assert np.abs(md.L - g["L1"]).max() < 1e-9
    assert np.abs(md.axes[:, 0:3] - g["a11"]).max() < 1e-9


def _rotate(ax9, axis, angle):
    """rigid rotation of one particle's axes about a lab-frame axis"""
    k = axis / np.linalg.norm(axis)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K
    return (ax9.reshape(3, 3) @ R.T).reshape(9)


@pytest.mark.parametrize("model", ["rna", "dna"])
def test_oracle_gradient_form_is_the_gradient(model):
    """With cpu_quirks = 0 the restated force and torque ARE minus the gradient of the restated energy (central differences).
    This is the form the CUDA kernels are held to (the reference's own CUDA kernels use the gradient as well)."""
    rng = np.random.default_rng(5)
    if model == "rna":
        g = load_golden("rna_lattice8_seqdep")
        P = _rna_params(g, cpu_quirks=False)
    else:
        g = load_golden("lattice8")
        P = _params(g)
    pos = g["pos"] + rng.normal(scale=0.01, size=g["pos"].shape)
    ax = O.axes_from_a1a3(g["a1"] + rng.normal(scale=0.05, size=g["a1"].shape), g["a3"] + rng.normal(scale=0.05, size=g["a3"].shape))
    pairs = O.verlet_pairs(pos, g["n3"], g["n5"], g["box"], P.rcut + 0.2)
    args = (g["btype"], g["n3"], g["n5"], g["box"], pairs)
    out = O.forces(P, pos, ax, *args)
    h = 1e-6
    worst_f = worst_t = 0.0
    for i in rng.choice(len(pos), size=24, replace=False):
        for k in range(3):
            e = np.zeros(3)
            e[k] = 1.0
            pp, pm = pos.copy(), pos.copy()
            pp[i] += h * e
            pm[i] -= h * e
            fd = -(O.forces(P, pp, ax, *args)["U"] - O.forces(P, pm, ax, *args)["U"]) / (2 * h)
            worst_f = max(worst_f, abs(fd - out["force"][i, k]))
            ap, am = ax.copy(), ax.copy()
            ap[i] = _rotate(ax[i], e, h)
            am[i] = _rotate(ax[i], e, -h)
            td = -(O.forces(P, pos, ap, *args)["U"] - O.forces(P, pos, am, *args)["U"]) / (2 * h)
            worst_t = max(worst_t, abs(td - out["torque_lab"][i, k]))
    scale = max(1.0, np.abs(out["force"]).max())
    assert worst_f < 2e-6 * scale and worst_t < 2e-6 * scale, (worst_f, worst_t, scale)


@pytest.mark.skipif(not RH.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_rna_oracle_matches_live_reference_without_meshed_term(tmp_path):
    """All-A sequence (no hydrogen bonding, hence no meshed factor): 60 strongly perturbed configurations of the reference's
    16-nt RNA test system, restatement (cpu_quirks = 1) against the live reference CPU class to rounding."""
    rng = np.random.default_rng(1)
    top = str(tmp_path / "aaaa.top")
    with open(top, "w") as f:
        f.write("16 3 5->3\nAAAA circular=False type=RNA\nAAAA circular=False type=RNA\nAAAAAAAA circular=False type=RNA\n")
    conf = os.path.join(GOLD, "force_field_rna", "init.dat")
    r = RH.Reference(top, conf, interaction_type="RNA2", salt_concentration=0.3, T="37C")
    try:
        st, topo = r.state(), r.topology()
        P = O.rna2_params(O.celsius(37.0), 0.3, cpu_quirks=True)
        assert P.rcut == r.rcut()
        worst = 0.0
        for it in range(60):
            sc = 0.05 * (it % 8)
            pos = st["pos"] + rng.normal(scale=sc / 2, size=st["pos"].shape)
            ax = O.axes_from_a1a3(st["a1"] + rng.normal(scale=sc, size=st["a1"].shape), st["a3"] + rng.normal(scale=sc, size=st["a3"].shape))
            r.set_state(pos, ax[:, 0:3], ax[:, 6:9])
            ref, es = r.compute_forces(), r.energy_split()
            pairs = O.verlet_pairs(pos, topo["n3"], topo["n5"], r.box(), P.rcut + 0.1)
            out = O.forces(P, pos, ax, topo["btype"], topo["n3"], topo["n5"], r.box(), pairs)
            worst = max(worst, np.abs(out["eterms"] - es).max() / max(1.0, np.abs(es).max()))
            for k in ("force", "torque_lab", "torque_body"):
                worst = max(worst, np.abs(out[k] - ref[k]).max() / max(1.0, np.abs(ref[k]).max()))
        assert worst < 1e-10, worst
    finally:
        r.close()


@pytest.mark.skipif(not RH.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_first_generation_oxrna_oracle_matches_live_reference():
    """interaction_type = RNA (no Debye-Hueckel): cutoff bit-equal, every term but the meshed hydrogen bonding to rounding"""
    top, conf = os.path.join(GOLD, "force_field_rna", "init.top"), os.path.join(GOLD, "force_field_rna", "init.dat")
    r = RH.Reference(top, conf, interaction_type="RNA", T="30C")
    try:
        P = O.rna2_params(O.celsius(30.0), 0.0, cpu_quirks=True)
        assert P.rcut == r.rcut()
        st, topo = r.state(), r.topology()
        r.compute_forces()
        es = r.energy_split()
        pairs = O.verlet_pairs(st["pos"], topo["n3"], topo["n5"], r.box(), P.rcut + 0.1)
        assert pair_set(pairs) == pair_set(r.pairs())
        out = O.forces(P, st["pos"], O.axes_from_a1a3(st["a1"], st["a3"]), topo["btype"], topo["n3"], topo["n5"], r.box(), pairs)
        d = out["eterms"][:7] - es[:7]
        assert np.abs(np.delete(d, 4)).max() < 1e-10 and abs(d[4]) < 1e-4 * abs(es[4]) and out["eterms"][7] == 0.0
    finally:
        r.close()


def test_first_generation_oxdna_oracle_matches_reference_fixture():
    """interaction_type = DNA_nomesh (class DNAInteraction): fixture written by the reference CPU backend"""
    g = load_golden("lattice8_dna1")
    P = O.dna1_params(parse_temperature(str(g["T"])))
    assert P.rcut == float(g["rcut"])
    ax = O.axes_from_a1a3(g["a1"], g["a3"])
    pairs = O.verlet_pairs(g["pos"], g["n3"], g["n5"], g["box"], P.rcut + 2 * 0.05)
    assert pair_set(pairs) == pair_set(g["pairs"])
    out = O.forces(P, g["pos"], ax, g["btype"], g["n3"], g["n5"], g["box"], pairs)
    assert np.abs(out["eterms"][:7] - g["energy_split"][:7]).max() < 1e-10
    # the fixture keeps a1 and a3 of the reference's (slightly non-orthonormal after 3,000 steps) orientation matrices; rebuilding
    # the axes from them moves the sites by ~1e-12, which the stiff excluded volume turns into ~1e-8 (the live test below feeds
    # identical axes to both sides and holds 1e-10)
    assert np.abs(out["force"] - g["force"]).max() < 1e-7 and np.abs(out["torque_lab"] - g["torque_lab"]).max() < 1e-7
    md = O.MD(P, g["pos"], ax, g["vel"], g["L"], g["btype"], g["n3"], g["n5"], g["box"], 0.003, 0.05)
    md.step(int(g["nve_steps"]))
    assert np.abs(md.pos - g["pos1"]).max() < 1e-8 and np.abs(md.vel - g["vel1"]).max() < 1e-7


@pytest.mark.skipif(not RH.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("grooving", [0, 1])
def test_first_generation_oxdna_oracle_matches_live_reference(grooving):
    """120 perturbed configurations of the reference's 16-nt DNA test system (most of them with coaxial stacking active:
    mirrored theta1 and the phi3 factor), with and without major_minor_grooving"""
    rng = np.random.default_rng(3)
    top, conf = os.path.join(GOLD, "force_field_dna", "init.top"), os.path.join(GOLD, "force_field_dna", "init.dat")
    r = RH.Reference(top, conf, interaction_type="DNA_nomesh", T="30C", major_minor_grooving=grooving, max_backbone_force=10)
    try:
        P = O.dna1_params(O.celsius(30.0), grooving=bool(grooving), max_backbone_force=10.0)
        assert P.rcut == r.rcut()
        st, topo = r.state(), r.topology()
        worst, coax = 0.0, 0
        for it in range(120):
            sc = 0.04 * (it % 8)
            pos = st["pos"] + rng.normal(scale=sc / 2, size=st["pos"].shape)
            ax = O.axes_from_a1a3(st["a1"] + rng.normal(scale=sc, size=st["a1"].shape), st["a3"] + rng.normal(scale=sc, size=st["a3"].shape))
            r.set_state(pos, ax[:, 0:3], ax[:, 6:9])
            ref, es = r.compute_forces(), r.energy_split()
            pairs = O.verlet_pairs(pos, topo["n3"], topo["n5"], r.box(), P.rcut + 0.1)
            out = O.forces(P, pos, ax, topo["btype"], topo["n3"], topo["n5"], r.box(), pairs)
            worst = max(worst, np.abs(out["eterms"][:7] - es[:7]).max() / max(1.0, np.abs(es).max()))
            for k in ("force", "torque_lab"):
                worst = max(worst, np.abs(out[k] - ref[k]).max() / max(1.0, np.abs(ref[k]).max()))
            coax += es[6] != 0
        assert worst < 1e-10 and coax > 50, (worst, coax)
    finally:
        r.close()


def test_oracle_barostat_rescale_properties():
    """volume-move restatement: the atomic move scales every coordinate, the molecular move translates each strand rigidly with its
    centre of mass (intra-strand distances unchanged, strand centres scaled); acceptance formula at dE = 0 reduces to the ideal-gas term"""
    g = load_golden("lattice8")
    box, new_box = np.array(g["box"], dtype=float), np.array(g["box"], dtype=float) * np.array([1.02, 0.97, 1.0])
    a = O.barostat_rescale(g["pos"], g["strand"], box, new_box, False)
    assert np.allclose(a, g["pos"] * new_box / box, rtol=0, atol=1e-14)
    m = O.barostat_rescale(g["pos"], g["strand"], box, new_box, True)
    for sid in np.unique(g["strand"]):
        sel = g["strand"] == sid
        assert np.allclose(m[sel] - m[sel].mean(0), g["pos"][sel] - g["pos"][sel].mean(0), atol=1e-12)
        assert np.allclose(m[sel].mean(0), g["pos"][sel].mean(0) * new_box / box, atol=1e-12)
    V0, V1 = box.prod(), new_box.prod()
    assert np.isclose(O.barostat_acceptance(0.0, 0.0, 0.1, box, new_box, 16), (V1 / V0) ** 16)


@pytest.mark.skipif(not RH.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("mode,extra", [("hb_cutoff", {}), ("switching_function", dict(d0=0.35, r0=0.45, n=6)),
                                        ("mixed", dict(mixed_weight=0.7, hb_energy_cutoff=-0.1, d0=0.4, r0=0.5, n=4))])
def test_meta_coordination_restatement_matches_live_reference(tmp_path, mode, extra):
    """The last external force of the reference's CUDA backend that the device side does not evaluate yet: its
    oracle is in place and pinned.  LTCoordination (smooth count of formed base pairs + tabulated bias) on the thermalised lattice8
    state: force and lab-frame torque on every particle of the 20 candidate pairs of duplex 0 against the unmodified reference CPU
    class, for the three coordination types."""
    g = load_golden("lattice8")
    top, conf, opf, ff = (str(tmp_path / n) for n in ("l.top", "l.dat", "op.txt", "forces.txt"))
    oio.write_topology(top, g["btype"], g["n3"], g["n5"], g["strand"])
    oio.write_conf(conf, g["box"], g["pos"], g["a1"], g["a3"], g["vel"], g["L"])
    pairs = [(k, 39 - k) for k in range(20)]
    with open(opf, "w") as f:
        f.write("{\norder_parameter = bond\nname = all_native_bonds\n" + "".join(f"pair{k + 1} = {a}, {b}\n" for k, (a, b) in enumerate(pairs)) + "}\n")
    xs = np.linspace(0.0, 20.2, 102)
    grid = 0.05 * (xs - 12.0) ** 2
    d = dict(coordination_type=mode, coord_min=0.0, coord_max=20.2, N_grid=len(grid), **extra)
    with open(ff, "w") as f:
        f.write("{\ntype = meta_coordination\nop_file = " + opf + "\n" + "".join(f"{k} = {v}\n" for k, v in d.items()) +
                "potential_grid = " + ",".join("%.10f" % v for v in grid) + "\n}\n")
    r = RH.Reference(top, conf, interaction_type="DNA2_nomesh", salt_concentration=float(g["salt"]), T=str(g["T"]), thermostat="no", dt=0.003,
                     external_forces=1, external_forces_file=ff)
    try:
        RH.lib().oxref_rebuild_lists()
        ref = r.compute_forces()
    finally:
        r.close()
    ax = O.axes_from_a1a3(g["a1"], g["a3"])
    out = O.meta_coordination(dict(d, pairs=pairs, potential_grid=[float("%.10f" % v) for v in grid]), g["pos"], ax, g["btype"], g["box"])
    dF, dT = ref["force"] - g["force"], ref["torque_lab"] - g["torque_lab"]
    assert 5.0 < out["coordination"] < 20.0
    assert np.abs(dF).max() > 1e-3 and np.abs(dT).max() > 1e-4          # the bias does act
    assert not np.abs(dF[40:]).max() > 1e-12                            # and only on the particles of the pairs
    assert np.abs(out["force"] - dF).max() < 1e-9
    assert np.abs(out["torque_lab"] - dT).max() < 1e-9


SEQ2 = "/root/reference/oxDNA2_sequence_dependent_parameters.txt"


@pytest.mark.skipif(not (RH.available() and os.path.exists(SEQ2)), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_dna2_sequence_dependent_with_dummy_bases_matches_live_reference(tmp_path):
    """oxDNA2 with the sequence-dependent stacking / hydrogen-bonding strengths and two dummy bases ('D': btype = type = 4,
    TopologyParser.cpp:96-99 -- they keep the average strengths, DNAInteraction.cpp:329-375, and pair with nothing)"""
    from oxdna_b200.sim import read_seq_dep
    g = load_golden("lattice8")
    bt = g["btype"].copy()
    bt[45], bt[130] = 4, 4
    top, conf = str(tmp_path / "t.top"), str(tmp_path / "t.dat")
    oio.write_topology(top, bt, g["n3"], g["n5"], g["strand"])
    oio.write_conf(conf, g["box"], g["pos"], g["a1"], g["a3"], g["vel"], g["L"])
    r = RH.Reference(top, conf, interaction_type="DNA2_nomesh", salt_concentration=0.5, T="300K", use_average_seq=0, seq_dep_file=SEQ2)
    try:
        ref, split, pairs, topo = r.compute_forces(), r.energy_split(), r.pairs(), r.topology()
    finally:
        r.close()
    assert (topo["btype"] == bt).all() and (topo["type"][[45, 130]] == 4).all()
    sd = read_seq_dep(SEQ2)
    B = "AGCT"
    P = O.dna2_params(parse_temperature("300K"), 0.5)
    O.dna2_params_seqdep(P, [sd[f"STCK_{a}_{b}"] for a in B for b in B], sd["STCK_FACT_EPS"], sd["HYDR_A_T"], sd["HYDR_C_G"])
    out = O.forces(P, g["pos"], O.axes_from_a1a3(g["a1"], g["a3"]), bt, g["n3"], g["n5"], g["box"], pairs)
    assert np.abs(out["eterms"] - split).max() < 1e-9
    assert np.abs(out["force"] - ref["force"]).max() < 1e-9
    assert np.abs(out["torque_body"] - ref["torque_body"]).max() < 1e-9


def test_reference_golden_vector_file_with_the_meshed_class():
    """test/DNA/FORCE_FIELD/AVG_SEQ/reference.dat was written by the MESHED class (interaction_type = DNA2).  With the reference's cubic meshes
    restated (P.mesh = 1: src/Interactions/Mesh.{h,cpp}, DNAInteraction.cpp:199-213) the per-term energies agree to the print precision of the
    file (6 decimals) -- the analytic form, which the CUDA kernels evaluate, is off in the 6th digit of the hydrogen bonding"""
    ref = np.loadtxt(os.path.join(GOLD, "force_field_dna", "reference_avg_seq.dat"))
    t = oio.read_topology(os.path.join(GOLD, "force_field_dna", "init.top"))
    c = oio.read_conf(os.path.join(GOLD, "force_field_dna", "init.dat"))
    P = O.dna2_params(O.celsius(20.0), 1.0)
    ax = O.axes_from_a1a3(c["a1"], c["a3"])
    pairs = O.verlet_pairs(c["pos"], t["n3"], t["n5"], c["box"], P.rcut + 0.1)
    analytic = O.forces(P, c["pos"], ax, t["btype"], t["n3"], t["n5"], c["box"], pairs)["eterms"] / t["N"]
    P.mesh = 1
    meshed = O.forces(P, c["pos"], ax, t["btype"], t["n3"], t["n5"], c["box"], pairs)["eterms"] / t["N"]
    assert np.abs(meshed - ref).max() <= 5.5e-7, (meshed, ref)
    assert np.abs(analytic - ref).max() > np.abs(meshed - ref).max()


@pytest.mark.skipif(not RH.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("itype", ["DNA2", "DNA"])
def test_oracle_with_meshes_matches_the_live_meshed_classes(tmp_path, itype):
    """interaction_type = DNA2 / DNA (the meshed CPU classes, the reference's defaults) on a perturbed dense configuration: with P.mesh = 1 the
    restatement reproduces forces, torques and per-term energies to rounding"""
    g = load_golden("lattice27_dense")
    rng = np.random.default_rng(21)
    N = len(g["pos"])
    pos = g["pos"] + rng.normal(0, 0.02, (N, 3))
    ax = O.axes_from_a1a3(g["a1"] + rng.normal(0, 0.05, (N, 3)), g["a3"] + rng.normal(0, 0.05, (N, 3)))
    top, conf = str(tmp_path / "t.top"), str(tmp_path / "t.dat")
    oio.write_topology(top, g["btype"], g["n3"], g["n5"], g["strand"])
    oio.write_conf(conf, g["box"], pos, ax[:, 0:3], ax[:, 6:9], g["vel"], g["L"])
    r = RH.Reference(top, conf, interaction_type=itype, salt_concentration=0.5, T="300K", max_backbone_force=10.0)
    try:
        st, ref, split, pairs = r.state(), r.compute_forces(), r.energy_split(), r.pairs()
    finally:
        r.close()
    T = parse_temperature("300K")
    P = O.dna2_params(T, 0.5, max_backbone_force=10.0) if itype == "DNA2" else O.dna1_params(T, max_backbone_force=10.0)
    ax = O.axes_from_a1a3(st["a1"], st["a3"])
    plain = O.forces(P, st["pos"], ax, g["btype"], g["n3"], g["n5"], g["box"], pairs)
    P.mesh = 1
    out = O.forces(P, st["pos"], ax, g["btype"], g["n3"], g["n5"], g["box"], pairs)
    n = len(split)
    assert np.abs(out["eterms"][:n] - split).max() < 1e-9
    # (first-generation oxDNA: 2e-8 on the two nucleotides of an end-to-end coaxial stack, 5 % of 4e-7 that the meshes change there -- the
    # cubic coefficients are difference quotients of node values and amplify their last bits; everything else to 1e-12)
    tol = 1e-9 if itype == "DNA2" else 1e-7
    assert np.abs(out["force"] - ref["force"]).max() < tol
    assert np.abs(out["torque_body"] - ref["torque_body"]).max() < tol
    # the analytic form (DNA2_nomesh, the CUDA kernels) differs from the meshed class by the interpolation error of its 6- to 250-interval
    # meshes: up to a few 1e-2 in force on a perturbed configuration
    assert 1e-9 < np.abs(plain["force"] - ref["force"]).max() < 0.2
