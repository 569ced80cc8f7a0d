"""CPU test of the product's FP32 pair-potential arithmetic (oxdna_b200/csrc/dna_model.cuh compiled for the host by nvcc)
against the double-precision oracle.  Tolerance: the north star's mixed-precision bound, |dF| <= 1e-5 * max|F|."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import oracle as O
from oxdna_b200 import capi
from oxdna_b200.sim import parse_temperature

SO = os.path.join(ROOT, "tests", "support", "_build", "libhostmodel.so")


@pytest.fixture(scope="module")
def hostlib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, "tests", "support", "host_model.cu")
    deps = [src] + [os.path.join(ROOT, "oxdna_b200", "csrc", f) for f in ("dna_model.cuh", "rna_model.cuh", "models.cuh", "common.cuh", "params.cpp", "dna3_model.cuh", "dna3_pack.h", "kernels.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-o", SO, src,
                               os.path.join(ROOT, "oxdna_b200", "csrc", "params.cpp")])
    return C.CDLL(SO)


@pytest.mark.parametrize("case", ["force_field_dna/ref_dna2_nomesh", "lattice8", "lattice27_dense"])
def test_fp32_formulation_within_mixed_tolerance(hostlib, case):
    g = load_golden(case)
    T, salt = parse_temperature(str(g["T"])), float(g["salt"])
    M, rcut = capi.dna2_params(T, salt)
    N = len(g["pos"])
    ax = np.ascontiguousarray(O.axes_from_a1a3(g["a1"], g["a3"]))
    pairs = np.ascontiguousarray(g["pairs"], dtype=np.int32)
    F, Tl, ep = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pos, box = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["box"], dtype=np.float64)
    bt, n3, n5 = (np.ascontiguousarray(g[k], dtype=np.int32) for k in ("btype", "n3", "n5"))
    hostlib.host_dna2_forces(C.byref(M), N, p(pos), p(ax), p(bt), p(n3), p(n5), p(box), p(pairs), C.c_longlong(len(pairs)), p(F), p(Tl), p(ep))
    fmax = np.linalg.norm(g["force"], axis=1).max()
    tmax = np.linalg.norm(g["torque_lab"], axis=1).max()
    assert np.linalg.norm(F - g["force"], axis=1).max() <= 1e-5 * fmax
    assert np.linalg.norm(Tl - g["torque_lab"], axis=1).max() <= 1e-5 * tmax
    assert abs(ep.sum() - float(g["U"])) <= 1e-6 * abs(float(g["U"]))


def rna_models(g):
    """(product FP32 block, oracle block in its gradient form) for an RNA fixture"""
    T, salt = parse_temperature(str(g["T"])), float(g["salt"])
    mis = float(g["mismatch"]) if "mismatch" in g else -1.0
    M, rcut = capi.rna2_params(T, salt, mismatch_repulsion=mis >= 0, mismatch_repulsion_strength=max(mis, 0.0))
    P = O.rna2_params(T, salt, cpu_quirks=False, mismatch_repulsion=mis >= 0, mismatch_repulsion_strength=max(mis, 0.0))
    if "sd_stck" in g:
        sd = [g["sd_stck"], float(g["sd_st_t_dep"]), g["sd_cross"], float(g["sd_hb_AT"]), float(g["sd_hb_GC"]), float(g["sd_hb_GT"])]
        capi.rna2_params_seqdep(M, T, *sd)
        O.rna2_params_seqdep(P, *sd)
    return M, rcut, P


def test_rna_parameter_block_matches_oracle():
    g = load_golden("rna_lattice8_seqdep")
    M, rcut, P = rna_models(g)
    assert rcut == P.rcut == float(g["rcut"])
    assert abs(M.dh_rc - P.dh_rc) < 1e-6 and abs(M.dh_b - P.dh_b) < 1e-7 * abs(P.dh_b) + 1e-9
    for i in range(4):
        for j in range(4):
            assert abs(M.stck_eps[5 * i + j] - P.stck.eps[i][j]) < 2e-7 and abs(M.hb_eps[5 * i + j] - P.hb.eps[i][j]) < 2e-7
            assert abs(M.stck_shift[5 * i + j] - P.stck.shift[i][j]) < 2e-7 and abs(M.crst_kfac[5 * i + j] - P.crst_kfac[i][j]) < 2e-7
    assert abs(M.mis_eps - P.mis_eps) < 2e-7 and abs(M.mis_shift - P.mis_shift) < 2e-7
    names = ("stck_t5 stck_t6 stck_tb1 stck_tb2 hb_t1 hb_t2 hb_t3 hb_t4 hb_t7 hb_t8 crst_t1 crst_t2 crst_t3 crst_t7 crst_t8 cxst_t1 cxst_t4 "
             "cxst_t5 cxst_t6").split()
    for k, n in enumerate(names):
        o = getattr(P, n)
        for f in "a b t0 ts tc".split():
            assert abs(getattr(M.f4[k], f) - getattr(o, f)) < 1e-6, (n, f)


@pytest.mark.parametrize("case", ["force_field_rna/ref_rna2", "force_field_rna/ref_rna2_seqdep", "rna_lattice8", "rna_lattice8_nohb", "rna_lattice8_seqdep"])
@pytest.mark.parametrize("perturb", [0.0, 0.06])
def test_rna_fp32_formulation_within_mixed_tolerance(hostlib, case, perturb):
    """oxRNA2: FP32 device functions against the restatement in its gradient form (cpu_quirks = 0, what the reference's CUDA
    kernels evaluate).  `perturb` shakes the orientations so that the rarely visited branches (phi factors, mirrored theta1,
    mismatch repulsion) are exercised as well."""
    g = load_golden(case)
    M, rcut, P = rna_models(g)
    N = len(g["pos"])
    rng = np.random.default_rng(11)
    ax = np.ascontiguousarray(O.axes_from_a1a3(g["a1"] + rng.normal(scale=perturb, size=g["a1"].shape), g["a3"] + rng.normal(scale=perturb, size=g["a3"].shape)))
    pos, box = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["box"], dtype=np.float64)
    bt, n3, n5 = (np.ascontiguousarray(g[k], dtype=np.int32) for k in ("btype", "n3", "n5"))
    pairs = np.ascontiguousarray(O.verlet_pairs(pos, n3, n5, box, P.rcut + 0.1), dtype=np.int32)
    ref = O.forces(P, pos, ax, bt, n3, n5, box, pairs)
    F, Tl, ep = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    hostlib.host_rna2_forces(C.byref(M), N, p(pos), p(ax), p(bt), p(n3), p(n5), p(box), p(pairs), C.c_longlong(len(pairs)), p(F), p(Tl), p(ep))
    fmax = np.linalg.norm(ref["force"], axis=1).max()
    tmax = np.linalg.norm(ref["torque_lab"], axis=1).max()
    # Thermalised fixtures: the north star's 1e-5.  Shaken (unphysical) orientations push bases into the quadratic smoothing
    # zone of the base-base excluded volume, whose stiffness 2 eps b = 1.6e4 turns the 4e-8 rounding of an FP32 orientation
    # into ~7e-4 of absolute force error whatever the formulation (the reference's float quaternions included): allow it.
    stiff = 1e-3 if perturb > 0 else 0.0
    assert np.linalg.norm(F - ref["force"], axis=1).max() <= 1e-5 * fmax + stiff
    assert np.linalg.norm(Tl - ref["torque_lab"], axis=1).max() <= 1e-5 * tmax + stiff
    assert abs(ep.sum() - ref["U"]) <= 2e-6 * abs(ref["U"]) + 1e-2 * stiff
    assert np.abs(ep - ref["epart"]).max() <= 1e-5 * max(1.0, np.abs(ref["epart"]).max()) + 1e-2 * stiff


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = 1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)
    return R  # columns = a1, a2, a3


def _pair_cloud(rng, n_pairs, site_p, site_q, dmin, dmax, spread=None):
    """n_pairs isolated two-particle systems on a grid: q placed so that |site_q(q) - site_p(p)| is uniform in (dmin, dmax).
    site_*: coefficients on (a1, a2, a3).  spread: if set, q's frame is p's frame times a random rotation of at most that angle."""
    Rp = _random_rotations(rng, n_pairs)
    if spread is None:
        Rq = _random_rotations(rng, n_pairs)
    else:
        ax = rng.normal(size=(n_pairs, 3))
        ax /= np.linalg.norm(ax, axis=1)[:, None]
        ang = rng.uniform(0, spread, size=n_pairs)
        K = np.zeros((n_pairs, 3, 3))
        K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 2], ax[:, 1], ax[:, 2], -ax[:, 0], -ax[:, 1], ax[:, 0]
        dR = np.eye(3)[None] + np.sin(ang)[:, None, None] * K + (1 - np.cos(ang))[:, None, None] * (K @ K)
        Rq = Rp @ dR
    u = rng.normal(size=(n_pairs, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    d = rng.uniform(dmin, dmax, size=n_pairs)
    side = int(np.ceil(n_pairs ** (1 / 3)))
    k = np.arange(n_pairs)
    grid = np.stack([k % side, (k // side) % side, k // (side * side)], axis=1) * 6.0 + 3.0
    sp = Rp @ np.asarray(site_p)
    sq = Rq @ np.asarray(site_q)
    xp = grid
    xq = grid + sp + u * d[:, None] - sq
    pos = np.empty((2 * n_pairs, 3))
    pos[0::2], pos[1::2] = xp, xq
    R = np.empty((2 * n_pairs, 3, 3))
    R[0::2], R[1::2] = Rp, Rq
    axes = np.ascontiguousarray(np.concatenate([R[:, :, 0], R[:, :, 1], R[:, :, 2]], axis=1))
    box = np.array([side * 6.0] * 3)
    return pos, axes, box


@pytest.mark.parametrize("kind", ["base", "stack", "bonded"])
def test_rna_fp32_pair_cloud_covers_rare_branches(hostlib, kind):
    """Isolated random pairs placed inside the radial window of (base) hydrogen bonding / mismatch repulsion / cross stacking,
    (stack) coaxial stacking incl. the phi3/phi4 triple products and the mirrored theta1, (bonded) 3'-5' stacking incl. the
    thetaB and phi factors.  Excluded volume is switched off in both models (it is the DNA code path, tested above, and its
    r^-13 wall would drown the angular terms in random overlaps)."""
    rng = np.random.default_rng({"base": 1, "stack": 2, "bonded": 3}[kind])
    T = parse_temperature("310K")
    M, rcut = capi.rna2_params(T, 0.5, mismatch_repulsion=True, mismatch_repulsion_strength=1.3)
    P = O.rna2_params(T, 0.5, cpu_quirks=False, mismatch_repulsion=True, mismatch_repulsion_strength=1.3)
    sd = [np.linspace(1.1, 1.7, 16), 1.97561, np.linspace(50.0, 70.0, 16), 0.82, 1.06, 0.51]
    capi.rna2_params_seqdep(M, T, *sd)
    O.rna2_params_seqdep(P, *sd)
    M.excl_eps = 0.0
    P.excl_eps = 0.0
    n = 60000
    if kind == "base":
        pos, axes, box = _pair_cloud(rng, n, (0.4, 0, 0), (0.4, 0, 0), 0.2, 0.8)
    elif kind == "stack":
        pos, axes, box = _pair_cloud(rng, n, (0.34, 0, 0), (0.34, 0, 0), 0.36, 0.64)
    else:
        pos, axes, box = _pair_cloud(rng, n, (-0.4, 0, 0.2), (-0.4, 0, 0.2), 0.761 - 0.2, 0.761 + 0.2, spread=1.2)
    N = 2 * n
    bt = rng.integers(0, 4, size=N).astype(np.int32)
    n3 = np.full(N, -1, dtype=np.int32)
    n5 = np.full(N, -1, dtype=np.int32)
    if kind == "bonded":
        n3[0::2] = np.arange(1, N, 2)
        n5[1::2] = np.arange(0, N, 2)
        pairs = np.zeros((0, 2), dtype=np.int32)
    else:
        pairs = np.ascontiguousarray(np.stack([np.arange(0, N, 2), np.arange(1, N, 2)], axis=1), dtype=np.int32)
    ref = O.forces(P, pos, axes, bt, n3, n5, box, pairs)
    F, Tl, ep = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    hostlib.host_rna2_forces(C.byref(M), N, p(pos), p(axes), p(bt), p(n3), p(n5), p(box), p(pairs), C.c_longlong(len(pairs)), p(F), p(Tl), p(ep))
    active = np.abs(ref["epart"]) > 1e-4
    # coverage: the sample must actually reach the branches it is meant for
    if kind == "base":
        assert active.sum() > 400
    elif kind == "stack":
        assert (np.abs(ref["epart"]) > 1e-3).sum() > 400 and ref["eterms"][6] < -1.0
    else:
        assert ref["eterms"][2] < -100.0
    fn, tn = np.linalg.norm(ref["force"], axis=1), np.linalg.norm(ref["torque_lab"], axis=1)
    scale = np.maximum(np.maximum(fn, tn), 1.0)
    # pair-wise bound against the pair's own force scale (floor 1, reduced units; cf. SURVEY 8(d) F_floor): 99.9 % of the
    # 120,000 particles inside 6e-6, the worst one inside 2e-5
    for got, want in ((F, ref["force"]), (Tl, ref["torque_lab"])):
        err = np.linalg.norm(got - want, axis=1) / scale
        assert np.quantile(err, 0.999) <= 6e-6 and err.max() <= 2e-5, (np.quantile(err, 0.999), err.max())
    assert (np.abs(ep - ref["epart"]) / np.maximum(np.abs(ref["epart"]), 1.0)).max() <= 1e-5


def test_first_generation_oxdna_fp32_formulation(hostlib):
    """interaction_type = DNA: FP32 device functions against the reference fixture (and, shaken, against the oracle)"""
    g = load_golden("lattice8_dna1")
    T = parse_temperature(str(g["T"]))
    M, rcut = capi.dna1_params(T)
    P = O.dna1_params(T)
    assert rcut == P.rcut == float(g["rcut"]) and M.v1 == 1
    N = len(g["pos"])
    rng = np.random.default_rng(2)
    pos, box = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["box"], dtype=np.float64)
    bt, n3, n5 = (np.ascontiguousarray(g[k], dtype=np.int32) for k in ("btype", "n3", "n5"))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for perturb in (0.0, 0.06):
        ax = np.ascontiguousarray(O.axes_from_a1a3(g["a1"] + rng.normal(scale=perturb, size=g["a1"].shape), g["a3"] + rng.normal(scale=perturb, size=g["a3"].shape)))
        pairs = np.ascontiguousarray(O.verlet_pairs(pos, n3, n5, box, P.rcut + 0.1), dtype=np.int32)
        ref = O.forces(P, pos, ax, bt, n3, n5, box, pairs)
        F, Tl, ep = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N)
        hostlib.host_dna2_forces(C.byref(M), N, p(pos), p(ax), p(bt), p(n3), p(n5), p(box), p(pairs), C.c_longlong(len(pairs)), p(F), p(Tl), p(ep))
        stiff = 1e-3 if perturb > 0 else 0.0
        assert np.linalg.norm(F - ref["force"], axis=1).max() <= 1e-5 * np.linalg.norm(ref["force"], axis=1).max() + stiff
        assert np.linalg.norm(Tl - ref["torque_lab"], axis=1).max() <= 1e-5 * np.linalg.norm(ref["torque_lab"], axis=1).max() + stiff
        assert abs(ep.sum() - ref["U"]) <= 2e-6 * abs(ref["U"]) + 1e-2 * stiff


def test_oxdna1_coaxial_pair_cloud(hostlib):
    """isolated random pairs inside the coaxial-stacking radial window: mirrored theta1 + f5(cos phi3)^2 of oxDNA"""
    rng = np.random.default_rng(7)
    T = parse_temperature("310K")
    M, rcut = capi.dna1_params(T)
    P = O.dna1_params(T)
    M.excl_eps = 0.0
    P.excl_eps = 0.0
    n = 60000
    pos, axes, box = _pair_cloud(rng, n, (0.34, 0, 0), (0.34, 0, 0), 0.19, 0.61)
    N = 2 * n
    bt = rng.integers(0, 4, size=N).astype(np.int32)
    none = np.full(N, -1, dtype=np.int32)
    pairs = np.ascontiguousarray(np.stack([np.arange(0, N, 2), np.arange(1, N, 2)], axis=1), dtype=np.int32)
    ref = O.forces(P, pos, axes, bt, none, none, box, pairs)
    F, Tl, ep = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    hostlib.host_dna2_forces(C.byref(M), N, p(pos), p(axes), p(bt), p(none), p(none), p(box), p(pairs), C.c_longlong(len(pairs)), p(F), p(Tl), p(ep))
    assert ref["eterms"][6] < -1.0 and (np.abs(ref["epart"]) > 1e-3).sum() > 200
    scale = np.maximum(np.maximum(np.linalg.norm(ref["force"], axis=1), np.linalg.norm(ref["torque_lab"], axis=1)), 1.0)
    for got, want in ((F, ref["force"]), (Tl, ref["torque_lab"])):
        err = np.linalg.norm(got - want, axis=1) / scale
        assert np.quantile(err, 0.999) <= 6e-6 and err.max() <= 2e-5, (np.quantile(err, 0.999), err.max())


@pytest.mark.parametrize("case", ["dna3_lattice8", "dna3_lattice27_dense"])
def test_dna3_fp32_formulation_within_mixed_tolerance(hostlib, case):
    """oxDNA3: the packed per-tetramer records + the FP32 device functions (csrc/dna3_model.cuh, dna3_pack.h) compiled for the host,
    against the fixture of the reference CPU class (DNA3Interaction_nomesh) and, with a nick that switches coaxial stacking on, the oracle"""
    g = load_golden(case)
    N = len(g["pos"])
    tab = np.ascontiguousarray(g["dna3_tables"], dtype=np.float64)
    S = capi.dna3_scalars(g["dna3_scalars"])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pos, box = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["box"], dtype=np.float64)
    ax = np.ascontiguousarray(O.axes_from_a1a3(g["a1"], g["a3"]))
    pairs = np.ascontiguousarray(g["pairs"], dtype=np.int32)
    bt = np.ascontiguousarray(g["btype"], dtype=np.int32)

    def run(n3, n5, pairs):
        n3, n5, pairs = (np.ascontiguousarray(x, dtype=np.int32) for x in (n3, n5, pairs))
        F, Tl, ep, es = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N), np.zeros(8)
        hostlib.host_dna3_forces(p(tab), C.byref(S), N, p(pos), p(ax), p(bt), p(n3), p(n5), p(box), p(pairs), C.c_longlong(len(pairs)), p(F), p(Tl),
                                 p(ep), p(es))
        return F, Tl, ep, es

    F, Tl, ep, es = run(g["n3"], g["n5"], pairs)
    fmax = np.linalg.norm(g["force"], axis=1).max()
    tmax = np.linalg.norm(g["torque_lab"], axis=1).max()
    assert np.linalg.norm(F - g["force"], axis=1).max() <= 1e-5 * fmax
    assert np.linalg.norm(Tl - g["torque_lab"], axis=1).max() <= 1e-5 * tmax
    assert abs(ep.sum() - float(g["U"])) <= 1e-6 * abs(float(g["U"]))
    assert np.abs(es - g["energy_split"]).max() <= 2e-6 * abs(float(g["U"]))
    # nicked strands: the stacked neighbours across the nick become a non-bonded pair inside the coaxial-stacking window
    n3, n5 = g["n3"].copy(), g["n5"].copy()
    starts = np.flatnonzero(g["n3"] < 0)
    cut = []
    for k, first in enumerate(starts[::3]):
        i = first + (1 if k % 3 == 0 else 9)
        j = n5[i]
        n5[i], n3[j] = -1, -1
        cut.append((min(i, j), max(i, j)))
    pairs2 = np.vstack([pairs, np.array(cut, dtype=np.int32)])
    P = O.dna3_params(g["dna3_tables"], g["dna3_scalars"])
    ref = O.forces(P, pos, ax, bt, n3, n5, box, pairs2)
    assert abs(ref["eterms"][6]) > 1e-2
    F, Tl, ep, es = run(n3, n5, pairs2)
    fmax = np.linalg.norm(ref["force"], axis=1).max()
    tmax = np.linalg.norm(ref["torque_lab"], axis=1).max()
    assert np.linalg.norm(F - ref["force"], axis=1).max() <= 1e-5 * fmax
    assert np.linalg.norm(Tl - ref["torque_lab"], axis=1).max() <= 1e-5 * tmax
    assert np.abs(es - ref["eterms"]).max() <= 2e-6 * abs(ref["U"])


@pytest.mark.skipif(not (os.path.exists("/root/reference/oxDNA3_sequence_dependent_parameters.txt") and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liboxref.so"))),
                    reason="needs the live reference (oracle/_ref) and its oxDNA3 parameter file")
def test_dna3_fp32_max_backbone_force_branch(hostlib, tmp_path):
    """oxDNA3 with max_backbone_force: tables (incl. _mbf_xmax_SD) from the live DNA3Interaction_nomesh, bonds stretched beyond xmax by a
    perturbation -- the far branch of the FENE term and its energy constant (dna3_pack.h) in the FP32 formulation against the oracle"""
    from oracle import refharness as RH
    from oxdna_b200 import io as oio
    g = load_golden("dna3_lattice8")
    rng = np.random.default_rng(4)
    N = len(g["pos"])
    pos = g["pos"] + rng.normal(0, 0.06, (N, 3))
    ax = O.axes_from_a1a3(g["a1"] + rng.normal(0, 0.08, (N, 3)), g["a3"] + rng.normal(0, 0.08, (N, 3)))
    top, conf = str(tmp_path / "t.top"), str(tmp_path / "t.dat")
    oio.write_topology(top, g["btype"], g["n3"], g["n5"], g["strand"])
    oio.write_conf(conf, g["box"], pos, ax[:, 0:3], ax[:, 6:9], g["vel"], g["L"])
    r = RH.Reference(top, conf, interaction_type="DNA3_nomesh", salt_concentration=0.5, T="300K", use_average_seq=0,
                     seq_dep_file="/root/reference/oxDNA3_sequence_dependent_parameters.txt", max_backbone_force=5.0)
    try:
        tab, sc = np.zeros((215, 900)), np.zeros(40)
        k = RH.lib().oxref_dna3_tables(RH._p(tab), RH._p(sc))
        st, pairs, box = r.state(), r.pairs(), r.box()
    finally:
        r.close()
    assert sc[1] == 1.0 and tab[3].max() > 0  # use_mbf, _mbf_xmax_SD
    P = O.dna3_params(tab, sc[:k])
    ax = np.ascontiguousarray(O.axes_from_a1a3(st["a1"], st["a3"]))
    ref = O.forces(P, st["pos"], ax, g["btype"], g["n3"], g["n5"], box, pairs)
    # some bonds are in the far branch
    back = st["pos"] + ax[:, 0:3] * P.back_a1 + ax[:, 3:6] * P.back_a2
    i = np.flatnonzero(g["n3"] >= 0)
    d = np.linalg.norm(back[g["n3"][i]] - back[i], axis=1)
    assert (np.abs(d - 0.7564) > tab[3].max()).sum() >= 3
    S = capi.dna3_scalars(sc[:k])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    posc, boxc = np.ascontiguousarray(st["pos"]), np.ascontiguousarray(box, dtype=np.float64)
    bt, n3, n5 = (np.ascontiguousarray(g[x], dtype=np.int32) for x in ("btype", "n3", "n5"))
    pr = np.ascontiguousarray(pairs, dtype=np.int32)
    F, Tl, ep, es = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N), np.zeros(8)
    tabc = np.ascontiguousarray(tab)
    hostlib.host_dna3_forces(p(tabc), C.byref(S), N, p(posc), p(ax), p(bt), p(n3), p(n5), p(boxc), p(pr), C.c_longlong(len(pr)), p(F), p(Tl), p(ep), p(es))
    fmax = np.linalg.norm(ref["force"], axis=1).max()
    assert np.linalg.norm(F - ref["force"], axis=1).max() <= 1e-4 * fmax
    assert abs(es[0] - ref["eterms"][0]) <= 1e-5 * abs(ref["eterms"][0])
    assert abs(ep.sum() - ref["U"]) <= 1e-5 * abs(ref["U"])


def test_dna3_fp32_special_base_types(hostlib):
    """dummy bases (btype = type = 4, sites at the oxDNA2 offsets) and a custom pair with hb_multiplier: FP32 device formulation against the
    oracle (which test_oracle_dna3.py pins to the live reference on the same construction)"""
    from conftest import dna3_special_types
    g = load_golden("dna3_lattice8")
    bt, sc = dna3_special_types(g)
    N = len(bt)
    P = O.dna3_params(g["dna3_tables"], sc)
    ax = np.ascontiguousarray(O.axes_from_a1a3(g["a1"], g["a3"]))
    ref = O.forces(P, g["pos"], ax, bt, g["n3"], g["n5"], g["box"], g["pairs"])
    assert abs(ref["eterms"][4] - float(g["energy_split"][4])) > 0.5
    S = capi.dna3_scalars(sc)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    tab = np.ascontiguousarray(g["dna3_tables"], dtype=np.float64)
    pos, box = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["box"], dtype=np.float64)
    btc, n3, n5 = (np.ascontiguousarray(x, dtype=np.int32) for x in (bt, g["n3"], g["n5"]))
    pairs = np.ascontiguousarray(g["pairs"], dtype=np.int32)
    F, Tl, ep, es = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N), np.zeros(8)
    hostlib.host_dna3_forces(p(tab), C.byref(S), N, p(pos), p(ax), p(btc), p(n3), p(n5), p(box), p(pairs), C.c_longlong(len(pairs)), p(F), p(Tl), p(ep), p(es))
    fmax, tmax = np.linalg.norm(ref["force"], axis=1).max(), np.linalg.norm(ref["torque_lab"], axis=1).max()
    # (pure FP32 here: the dummy bases put excluded-volume site pairs in range, which the device re-evaluates in double -- DESIGN 4;
    # the GPU test of the same construction holds 1e-5)
    assert np.linalg.norm(F - ref["force"], axis=1).max() <= 3e-5 * fmax
    assert np.linalg.norm(Tl - ref["torque_lab"], axis=1).max() <= 3e-5 * tmax
    assert np.abs(es - ref["eterms"]).max() <= 2e-6 * abs(ref["U"])
