"""TEST INFRASTRUCTURE ONLY.  ctypes view of the C restatement oracle/oxdna_oracle.c (built with gcc into
oracle/_build/liboxoracle.so).  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboxoracle.so")
_SRC = [os.path.join(_HERE, "oxdna_oracle.c")]
_HDR = [os.path.join(_HERE, "oxdna_oracle.h"), os.path.join(_HERE, "oxrna_oracle.inc"), os.path.join(_HERE, "oxdna3_oracle.inc")]

NTERMS = 8
TERM_NAMES = ["FENE", "BEXC", "STCK", "NEXC", "HB", "CRSTCK", "CXSTCK", "DH"]


def build(force=False):
    """gcc -O2 (no -ffast-math: the oracle is plain IEEE double)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    newest = max(os.path.getmtime(p) for p in _SRC + _HDR)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", _SO] + _SRC + ["-lm"]
        subprocess.check_call(cmd)
    return _SO


class _F1(C.Structure):
    _fields_ = [(n, C.c_double) for n in "a rc r0 blow bhigh rlow rhigh rclow rchigh".split()] + [
        ("eps", (C.c_double * 5) * 5), ("shift", (C.c_double * 5) * 5)]


class _F2(C.Structure):
    _fields_ = [(n, C.c_double) for n in "k rc r0 blow rlow rclow bhigh rhigh rchigh".split()]


class _F4(C.Structure):
    _fields_ = [(n, C.c_double) for n in "a b t0 ts tc".split()]


class _F5(C.Structure):
    _fields_ = [(n, C.c_double) for n in "a b xc xs".split()]


class _Excl(C.Structure):
    _fields_ = [(n, C.c_double) for n in "sigma rstar b rc".split()]


class DNA2Params(C.Structure):
    _fields_ = (
        [(n, C.c_double) for n in "back_a1 back_a2 stack_a1 base_a1 backref_a1 T fene_eps fene_r0 fene_delta fene_delta2".split()]
        + [("use_mbf", C.c_int)]
        + [(n, C.c_double) for n in "mbf_xmax mbf_fmax mbf_finf".split()]
        + [("excl", _Excl * 4), ("excl_eps", C.c_double), ("hb", _F1), ("stck", _F1), ("crst", _F2), ("cxst", _F2)]
        + [(n, _F4) for n in "stck_t4 stck_t5 hb_t1 hb_t2 hb_t4 hb_t7 crst_t1 crst_t2 crst_t4 crst_t7 cxst_t1 cxst_t4 cxst_t5".split()]
        + [("cxst_t1_sa", C.c_double), ("cxst_t1_sb", C.c_double), ("stck_phi1", _F5), ("stck_phi2", _F5)]
        + [(n, C.c_double) for n in "dh_minus_kappa dh_prefactor dh_rhigh dh_rc dh_b".split()]
        + [("dh_half_charged_ends", C.c_int), ("hb_multiplier", C.c_double), ("rcut", C.c_double), ("v1", C.c_int), ("cxst_phi3", _F5), ("mesh", C.c_int)]
    )


class RNA2Params(C.Structure):
    _fields_ = (
        [("back", C.c_double * 3), ("stack_a1", C.c_double), ("base_a1", C.c_double), ("stack3", C.c_double * 2),
         ("stack5", C.c_double * 2), ("p3", C.c_double * 3), ("p5", C.c_double * 3)]
        + [(n, C.c_double) for n in "T fene_eps fene_r0 fene_delta fene_delta2".split()]
        + [("use_mbf", C.c_int)]
        + [(n, C.c_double) for n in "mbf_xmax mbf_fmax mbf_finf".split()]
        + [("excl", _Excl * 4), ("excl_eps", C.c_double), ("hb", _F1), ("stck", _F1), ("crst", _F2), ("cxst", _F2),
           ("crst_kfac", (C.c_double * 5) * 5)]
        + [(n, _F4) for n in ("stck_t5 stck_t6 stck_tb1 stck_tb2 hb_t1 hb_t2 hb_t3 hb_t4 hb_t7 hb_t8 crst_t1 crst_t2 crst_t3 crst_t7 "
                              "crst_t8 cxst_t1 cxst_t4 cxst_t5 cxst_t6").split()]
        + [(n, _F5) for n in "stck_phi1 stck_phi2 cxst_phi3 cxst_phi4".split()]
        + [(n, C.c_double) for n in "dh_minus_kappa dh_prefactor dh_rhigh dh_rc dh_b".split()]
        + [("dh_half_charged_ends", C.c_int), ("average", C.c_int), ("mismatch_repulsion", C.c_int), ("mis_eps", C.c_double),
           ("mis_shift", C.c_double), ("cpu_quirks", C.c_int), ("rcut", C.c_double)]
    )


class DNA3Params(C.Structure):
    """oxDNA3: tetramer-indexed tables (a pointer into `tables`, kept alive by the wrapper) + scalars; see oxdna_oracle.h"""
    _fields_ = ([("tab", C.c_void_p), ("fene_eps", C.c_double), ("use_mbf", C.c_int)]
                + [(n, C.c_double) for n in "mbf_fmax mbf_finf hb_multiplier dh_rc dh_rhigh dh_prefactor dh_b dh_minus_kappa".split()]
                + [("dh_half_charged_ends", C.c_int), ("rcut", C.c_double), ("cxst_t1", _F4), ("cxst_t4", _F4), ("cxst_t5", _F4)]
                + [(n, C.c_double) for n in "cxst_t1_sa cxst_t1_sb excl_eps back_a1 back_a2 backref_a1".split()]
                + [("pos_stack", C.c_double * 5), ("pos_base", C.c_double * 5), ("ref_form", C.c_int), ("cxst_mesh", C.c_int)])


DNA3_NTAB, DNA3_TSIZE, DNA3_NSCALARS = 215, 900, 29


class ExtForce(C.Structure):
    _fields_ = [("type", C.c_int), ("particle", C.c_int), ("ref", C.c_int), ("pbc", C.c_int)] + [
        (n, C.c_double) for n in "stiff r0 rate stiff_rate F0".split()] + [("dir", C.c_double * 3), ("pos0", C.c_double * 3),
                                                                              ("aux", C.c_double * 8), ("iaux", C.c_int)]

EXT_TYPES = {"string": 0, "trap": 1, "mutual_trap": 2, "lowdim_trap": 3, "repulsion_plane": 4, "attraction_plane": 5, "sphere": 6, "LJ_wall": 7, "twist": 8, "sphere_smooth": 9, "ellipsoid": 10, "repulsion_plane_moving": 11, "generic_central_force": 12, "LJ_cone": 13, "com": 14, "yukawa_sphere": 15, "repulsive_sphere_moving": 16, "meta_com_trap": 17, "meta_coordination": 18}


def _index_list(v):
    """particle lists of the forces file: an int, a sequence, or the reference's "a,b,c" / "a-b" strings (Utils::get_particles_from_string)"""
    if isinstance(v, (int, np.integer)):
        return [int(v)]
    if isinstance(v, str):
        out = []
        for tok in v.split(","):
            tok = tok.strip()
            if "-" in tok[1:]:
                a, b = tok.split("-")
                out.extend(range(int(a), int(b) + 1))
            else:
                out.append(int(tok))
        return out
    return [int(x) for x in v]


def fill_ext_entry(e, d, pool, grid):
    """dict with the reference's external-force keys (docs/source/forces.md) -> one table entry (oxb_ext_force / oxo_ext_force)"""
    e.type = EXT_TYPES[d["type"]]
    part = d.get("particle", -1)
    e.particle = -1 if str(part) in ("-1", "all") else int(part)
    e.ref = int(d.get("ref_particle", -1)) if d["type"] != "repulsion_plane_moving" else -1
    e.pbc = int(d.get("PBC", 0)) if d["type"] not in ("meta_com_trap", "meta_coordination") else 0
    e.stiff, e.r0, e.rate = float(d.get("stiff", 1.0 if d["type"] == "LJ_wall" else 0.0)), float(d.get("r0", 0.0)), float(d.get("rate", 0.0))
    e.stiff_rate, e.F0 = float(d.get("stiff_rate", 0.0)), float(d.get("F0", 0.0))
    dr = np.array(d.get("axis", d.get("dir", (1, 0, 0) if d["type"] != "mutual_trap" else (0, 0, 1))), dtype=np.float64)
    if d["type"] != "mutual_trap" and np.linalg.norm(dr) > 0:
        dr = dr / np.linalg.norm(dr)
    centre = d.get("pos0", (0, 0, 0)) if d["type"] == "twist" else d.get("center", d.get("pos0", (0, 0, 0)))
    if d["type"] == "string" and int(d.get("dir_as_centre", 0)):
        # ConstantRateForce with dir_as_centre (src/Forces/ConstantRateForce.cpp:31,44-46,54-61): `dir` is a point, kept unnormalised in
        # pos0; the force points from the particle towards it.  Flagged in the (otherwise unused) pbc field.
        centre = d["dir"]
        e.pbc = 1
    for c in range(3):
        e.dir[c] = dr[c]
        e.pos0[c] = float(centre[c])
    aux = [0.0] * 8
    e.iaux = 0
    if d["type"] == "lowdim_trap":
        vis = d.get("visibility", (1, 1, 1))
        e.iaux = (1 if vis[0] else 0) | (2 if vis[1] else 0) | (4 if vis[2] else 0)
    elif d["type"] == "repulsion_plane":
        aux[0], aux[1], aux[2] = float(d["position"]), float(d.get("v", 0.0)), float(d.get("end_position", 1e6))
    elif d["type"] == "attraction_plane":
        aux[0] = float(d["position"])
    elif d["type"] == "sphere":
        aux[0] = float(d.get("r_ext", 1e10))
    elif d["type"] == "LJ_wall":
        n = int(d.get("n", 6))
        aux[0], aux[1] = float(d["position"]), float(d.get("sigma", 1.0))
        aux[2] = 2.0 ** (1.0 / n) if int(d.get("only_repulsive", 0)) else 1e6
        e.iaux = n
    elif d["type"] == "twist":
        e.F0 = float(d.get("base", 0.0))
        aux[0:3] = [float(x) for x in d["center"]]
        aux[3:6] = [float(x) for x in d.get("mask", (0.0, 0.0, 0.0))]
    elif d["type"] == "sphere_smooth":
        # the reference reads `smooth` and `alpha` from the r_ext key as well (RepulsiveSphereSmooth.cpp:27-29)
        aux[0] = float(d["r_ext"])
        aux[1], aux[2] = float(d.get("smooth", aux[0])), float(d.get("alpha", aux[0]))
    elif d["type"] == "ellipsoid":
        aux[0:3] = [float(x) for x in d["r_2"]]
        aux[3:6] = [float(x) for x in d.get("r_1", (1e-6, 1e-6, 1e-6))]
    elif d["type"] == "repulsion_plane_moving":
        refs = sorted(_index_list(d["ref_particle"]))
        if refs[-1] - refs[0] + 1 != len(refs):
            raise ValueError("RepulsionPlaneMoving requires the list of ref_particle indices to be contiguous")
        e.ref, e.iaux = refs[0], refs[-1]
    elif d["type"] == "generic_central_force":
        if d.get("force_type", "gravity") != "gravity":
            raise ValueError("only force_type = gravity runs on the device (as in the reference's CUDA backend)")
        aux[0], aux[1] = float(d.get("inner_cut_off", 0.0)) ** 2, float(d.get("outer_cut_off", 0.0)) ** 2
    elif d["type"] == "LJ_cone":
        n = int(d.get("n", 6))
        e.stiff = float(d.get("stiff", 1.0))
        aux[0], aux[2] = float(d.get("sigma", 1.0)), float(d["alpha"])
        aux[1] = 2.0 ** (1.0 / n) if int(d.get("only_repulsive", 0)) else 1e6
        e.iaux = n
    elif d["type"] == "com":
        com, ref = _index_list(d["com_list"]), _index_list(d["ref_list"])
        # COMForce keeps std::set<BaseParticle *> lists: duplicates collapse
        com, ref = sorted(set(com)), sorted(set(ref))
        e.particle, e.ref, e.iaux, e.pbc = -1, len(pool), len(com), len(ref)
        pool.extend(com + ref)
    elif d["type"] == "yukawa_sphere":
        # YukawaSphere.cpp:22-25 reads WCA_n into sigma (sic); the exponent stays at its default of 6
        sigma = float(d.get("WCA_n", d.get("WCA_sigma", 1.0)))
        e.r0, e.stiff, e.iaux = float(d["radius"]), float(d.get("WCA_epsilon", 1.0)), 6
        aux[0], aux[1] = sigma, sigma * 2.0 ** (1.0 / 6)
        aux[2], aux[3] = float(d["debye_length"]), float(d["debye_A"])
        aux[4] = float(d.get("cutoff", 4.0 * aux[2]))
    elif d["type"] == "repulsive_sphere_moving":
        org = d.get("origin", d.get("center", (0.0, 0.0, 0.0)))
        for c in range(3):
            e.pos0[c] = float(org[c])
        aux[0] = float(d.get("r_ext", 1e10))
        aux[1:4] = [float(x) for x in d.get("target", (0.0, 0.0, 0.0))]
        aux[4] = float(int(float(d.get("steps", d.get("move_steps", 0)))))
    elif d["type"] == "meta_com_trap":
        p1a, p2a = _index_list(d["p1a"]), _index_list(d["p2a"])
        pg = d["potential_grid"]
        pg = [float(x) for x in (pg.split(",") if isinstance(pg, str) else pg)]
        n_grid = int(d["N_grid"])
        if len(pg) != n_grid:
            raise ValueError("meta_com_trap: potential_grid must hold N_grid values")
        e.particle, e.ref, e.iaux, e.pbc = -1, len(pool), len(p1a), len(p2a)
        pool.extend(p1a + p2a)
        aux[0], aux[1], aux[2] = float(d["xmin"]), (float(d["xmax"]) - float(d["xmin"])) / (n_grid - 1.0), float(n_grid)
        aux[3], aux[4], aux[5] = float(int(d["mode"])), float(len(grid)), float(int(d.get("PBC", 0)))
        e.pbc = len(p2a)
        grid.extend(pg)
    elif d["type"] == "meta_coordination":
        # `pairs`: the hydrogen-bond candidate pairs of the op_file (LTCoordination::init reads them from there)
        pairs = [(int(a), int(b)) for a, b in d["pairs"]]
        flat = [x for ab in pairs for x in ab]
        if len(set(flat)) != len(flat):
            raise ValueError("LTCoordination assumes each particle appears only once")
        pg = d["potential_grid"]
        pg = [float(x) for x in (pg.split(",") if isinstance(pg, str) else pg)]
        n_grid = int(d["N_grid"])
        if len(pg) != n_grid:
            raise ValueError("LTCoordination: potential_grid size != N_grid")
        mode = {"hb_cutoff": 0, "switching_function": 1, "mixed": 2}[d.get("coordination_type", "hb_cutoff")]
        cmin, cmax = float(d.get("coord_min", 0.0)), float(d.get("coord_max", len(pairs) * 1.01))
        e.particle, e.ref, e.iaux, e.pbc = -1, len(pool), len(pairs), int(d.get("n", 6))
        pool.extend(flat)
        e.r0, e.stiff, e.F0 = float(d.get("d0", 0.4)), float(d.get("r0", 0.5)), cmax
        aux[0], aux[1], aux[2], aux[3], aux[4] = cmin, (cmax - cmin) / (n_grid - 1.0), float(n_grid), float(mode), float(len(grid))
        aux[5], aux[6], aux[7] = float(d.get("mixed_weight", 0.0)), float(d.get("hb_energy_cutoff", -0.2)), float(d.get("hb_transition_width", 0.1))
        grid.extend(pg)
    for c in range(8):
        e.aux[c] = aux[c]


EXT_STRING, EXT_TRAP, EXT_MUTUAL = 0, 1, 2


class _MD(C.Structure):
    _fields_ = [("N", C.c_int)] + [(n, C.c_void_p) for n in "pos axes vel L force torque_body list_pos btype n3 n5".split()] + [
        ("box", C.c_double * 3), ("dt", C.c_double), ("skin", C.c_double), ("pairs", C.c_void_p), ("npairs", C.c_longlong),
        ("max_pairs", C.c_longlong), ("step", C.c_longlong), ("nf", C.c_int), ("ef", C.c_void_p), ("U", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oxo_verlet_pairs.restype = C.c_longlong
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def kelvin(T):
    """src/Utilities/Utils.cpp:316-346"""
    return T * 0.1 / 300.0


def celsius(T):
    return (T + 273.15) * 0.1 / 300.0


def dna2_params(T, salt=0.5, dh_half_charged_ends=True, max_backbone_force=None, max_backbone_force_far=0.04):
    P = DNA2Params()
    mbf = max_backbone_force is not None
    lib().oxo_dna2_params_init(C.byref(P), C.c_double(T), C.c_double(salt), int(dh_half_charged_ends), int(mbf),
                               C.c_double(max_backbone_force if mbf else 0.0),
                               C.c_double(float(np.float32(max_backbone_force_far))))
    return P


def dna1_params(T, grooving=False, max_backbone_force=None, max_backbone_force_far=0.04):
    """first-generation oxDNA (interaction_type = DNA)"""
    P = DNA2Params()
    mbf = max_backbone_force is not None
    lib().oxo_dna1_params_init(C.byref(P), C.c_double(T), int(grooving), int(mbf), C.c_double(max_backbone_force if mbf else 0.0),
                               C.c_double(float(np.float32(max_backbone_force_far))))
    return P


def dna2_params_seqdep(P, stck16, stck_fact_eps, hb_AT, hb_GC):
    s = _d(stck16).reshape(16)
    lib().oxo_dna2_params_seqdep(C.byref(P), _p(s), C.c_double(stck_fact_eps), C.c_double(hb_AT), C.c_double(hb_GC))
    return P


def rna2_params(T, salt=1.0, dh_half_charged_ends=True, max_backbone_force=None, max_backbone_force_far=0.04,
                mismatch_repulsion=False, mismatch_repulsion_strength=1.0, cpu_quirks=False):
    """oxRNA2 parameters.  cpu_quirks=True (= 3; 1 and 2 select one of them) reproduces the two spots where the reference CPU class's
    force is not the gradient of its energy (see oxdna_oracle.h); the reference's CUDA kernels -- and ours -- use the gradient."""
    P = RNA2Params()
    mbf = max_backbone_force is not None
    lib().oxo_rna2_params_init(C.byref(P), C.c_double(T), C.c_double(salt), int(dh_half_charged_ends), int(mbf),
                               C.c_double(max_backbone_force if mbf else 0.0),
                               C.c_double(float(np.float32(max_backbone_force_far))), int(mismatch_repulsion),
                               C.c_double(mismatch_repulsion_strength))
    P.cpu_quirks = 3 if cpu_quirks is True else int(cpu_quirks)  # bit 0: phi2 stacking term, bit 1: mirrored coaxial theta1 term
    return P


def rna2_params_seqdep(P, stck16, st_t_dep, cross16, hb_AT, hb_GC, hb_GT):
    s, c = _d(stck16).reshape(16), _d(cross16).reshape(16)
    lib().oxo_rna2_params_seqdep(C.byref(P), _p(s), C.c_double(st_t_dep), _p(c), C.c_double(hb_AT), C.c_double(hb_GC), C.c_double(hb_GT))
    return P


def dna3_params(tables, scalars):
    """tables: (215, 900) doubles, scalars: the block of oxref_dna3_tables -- both as stored in tests/golden/dna3_*.npz"""
    P = DNA3Params()
    P._tables = np.ascontiguousarray(tables, dtype=np.float64).reshape(DNA3_NTAB, DNA3_TSIZE)
    P._scalars = np.ascontiguousarray(scalars, dtype=np.float64)
    lib().oxo_dna3_params_fill(C.byref(P), _p(P._tables), _p(P._scalars))
    return P


def axes_from_a1a3(a1, a3):
    a1, a3 = _d(a1), _d(a3)
    N = a1.shape[0]
    ax = np.zeros((N, 9))
    lib().oxo_axes_from_a1a3(N, _p(a1), _p(a3), _p(ax))
    return ax


def verlet_pairs(pos, n3, n5, box, rv):
    pos, n3, n5, box = _d(pos), _i(n3), _i(n5), _d(box)
    N = pos.shape[0]
    cap = max(64 * N, 1024)
    while True:
        out = np.zeros((cap, 2), dtype=np.int32)
        n = lib().oxo_verlet_pairs(N, _p(pos), _p(n3), _p(n5), _p(box), C.c_double(rv), _p(out), C.c_longlong(cap))
        if n <= cap:
            return out[:n].copy()
        cap = n


def forces(P, pos, axes, btype, n3, n5, box, pairs):
    pos, axes, box = _d(pos), _d(axes), _d(box)
    btype, n3, n5, pairs = _i(btype), _i(n3), _i(n5), _i(pairs)
    N = pos.shape[0]
    f, tl, tb = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
    et, ep = np.zeros(NTERMS), np.zeros(N)
    fn = lib().oxo_rna2_forces if isinstance(P, RNA2Params) else (lib().oxo_dna3_forces if isinstance(P, DNA3Params) else lib().oxo_dna2_forces)
    fn(C.byref(P), N, _p(pos), _p(axes), _p(btype), _p(n3), _p(n5), _p(box), _p(pairs),
                          C.c_longlong(pairs.shape[0]), _p(f), _p(tl), _p(tb), _p(et), _p(ep))
    return dict(force=f, torque_lab=tl, torque_body=tb, eterms=et, epart=ep, U=et.sum())


_pool_keep = None


def make_ext(forces_list):
    """table entries; the index pool of the COM forces is handed to the library here (kept alive at module level)"""
    global _pool_keep
    arr = (ExtForce * max(len(forces_list), 1))()
    pool, grid = [], []
    for k, d in enumerate(forces_list):
        fill_ext_entry(arr[k], d, pool, grid)
    _pool_keep = ((C.c_int * max(len(pool), 1))(*pool), (C.c_double * max(len(grid), 1))(*grid))
    lib().oxo_set_ext_pool(_pool_keep[0])
    lib().oxo_set_ext_grid(_pool_keep[1])
    return arr


def ext_forces(ext, pos, box, step):
    pos, box = _d(pos), _d(box)
    f = np.zeros_like(pos)
    arr = make_ext(ext)
    lib().oxo_ext_forces(len(ext), arr, pos.shape[0], _p(pos), _p(box), C.c_longlong(step), _p(f))
    return f


class MD:
    """NVE velocity-Verlet driver around oxo_md_steps (MD_CPUBackend restatement)."""

    def __init__(self, P, pos, axes, vel, L, btype, n3, n5, box, dt, skin, ext=()):
        self.P = P
        self.N = N = pos.shape[0]
        self.pos, self.axes, self.vel, self.L = _d(pos).copy(), _d(axes).copy(), _d(vel).copy(), _d(L).copy()
        self.force, self.torque_body, self.list_pos = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
        self.btype, self.n3, self.n5 = _i(btype).copy(), _i(n3).copy(), _i(n5).copy()
        self.max_pairs = 80 * N + 1024
        self.pairs = np.zeros((self.max_pairs, 2), dtype=np.int32)
        self.ext = make_ext(list(ext))
        S = self.S = _MD()
        S.N = N
        for n in "pos axes vel L force torque_body list_pos btype n3 n5 pairs".split():
            setattr(S, n, getattr(self, n).ctypes.data)
        for k in range(3):
            S.box[k] = float(box[k])
        S.dt, S.skin, S.max_pairs, S.step = dt, skin, self.max_pairs, 0
        S.nf = len(ext)
        S.ef = C.cast(self.ext, C.c_void_p)
        p = verlet_pairs(self.pos, self.n3, self.n5, box, P.rcut + 2 * skin)
        self.pairs[: len(p)] = p
        S.npairs = len(p)
        self.list_pos[:] = self.pos
        kind = "rna2_" if isinstance(P, RNA2Params) else ("dna3_" if isinstance(P, DNA3Params) else "")
        self._steps = getattr(lib(), f"oxo_{kind}md_steps")
        getattr(lib(), f"oxo_{kind}md_compute_forces")(C.byref(P), C.byref(S))

    def step(self, n=1):
        return self._steps(C.byref(self.P), C.byref(self.S), int(n))

    @property
    def U(self):
        return self.S.U

    def current_pairs(self):
        return self.pairs[: self.S.npairs].copy()


def brownian_params(T, dt, newtonian_steps, pt=0.0, diff_coeff=0.0):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    lib().oxo_brownian_params(C.c_double(T), C.c_double(dt), int(newtonian_steps), C.c_double(pt), C.c_double(diff_coeff),
                              C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def langevin_params(T, dt, gamma_trans=0.0, diff_coeff=0.0):
    v = [C.c_double() for _ in range(4)]
    lib().oxo_langevin_params(C.c_double(T), C.c_double(dt), C.c_double(gamma_trans), C.c_double(diff_coeff), *[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def barostat_rescale(pos, strand, box, new_box, molecular):
    """positions after a volume move: VolumeMove::apply (src/Backends/MCMoves/VolumeMove.cpp:79-90, every position scaled) or
    MoleculeVolumeMove::apply (src/Backends/MCMoves/MoleculeVolumeMove.cpp:78-104, every strand translated by com * (L'/L - 1));
    same arithmetic as the CUDA kernels rescale_positions / rescale_molecular_positions (src/CUDA/Backends/CUDA_MD.cuh:62-95)."""
    pos, box, new_box = _d(pos), _d(box), _d(new_box)
    if not molecular:
        return pos * (new_box / box)
    out = pos.copy()
    strand = np.asarray(strand)
    for sid in np.unique(strand):
        m = strand == sid
        out[m] += pos[m].mean(axis=0) * (new_box / box - 1.0)
    return out


def barostat_acceptance(dE, P, T, box, new_box, n_objs):
    """VolumeMove.cpp:98-102 / MD_CUDABackend.cu:488-492: exp(-(dE + P dV - N_objs T ln(V'/V)) / T)"""
    V0, V1 = float(np.prod(_d(box))), float(np.prod(_d(new_box)))
    return float(np.exp(-(dE + P * (V1 - V0) - n_objs * T * np.log(V1 / V0)) / T))


class Coord(C.Structure):
    _fields_ = [("mode", C.c_int), ("mixed_weight", C.c_double), ("hb_energy_cutoff", C.c_double), ("hb_transition_width", C.c_double),
                ("d0", C.c_double), ("r0", C.c_double), ("n", C.c_int), ("coord_min", C.c_double), ("coord_max", C.c_double), ("N_grid", C.c_int),
                ("grid", C.POINTER(C.c_double)), ("n_pairs", C.c_int), ("pairs", C.POINTER(C.c_int))]


def meta_coordination(d, pos, axes, btype, box):
    """meta_coordination force of the forces file (keys of LTCoordination::init and CoordSettings::get_settings; `pairs` = the hydrogen-bond
    pairs of the op_file).  Returns dict(coordination, force, torque_lab)."""
    pairs = _i(np.asarray(d["pairs"]).reshape(-1, 2))
    grid = _d(d["potential_grid"])
    c = Coord()
    c.mode = {"hb_cutoff": 0, "switching_function": 1, "mixed": 2}[d.get("coordination_type", "hb_cutoff")]
    c.mixed_weight = float(d.get("mixed_weight", 0.0))
    c.hb_energy_cutoff, c.hb_transition_width = float(d.get("hb_energy_cutoff", -0.2)), float(d.get("hb_transition_width", 0.1))
    c.d0, c.r0, c.n = float(d.get("d0", 0.4)), float(d.get("r0", 0.5)), int(d.get("n", 6))
    c.coord_min, c.coord_max = float(d.get("coord_min", 0.0)), float(d.get("coord_max", len(pairs) * 1.01))
    c.N_grid = len(grid)
    c.grid = grid.ctypes.data_as(C.POINTER(C.c_double))
    c.n_pairs = len(pairs)
    c.pairs = pairs.ctypes.data_as(C.POINTER(C.c_int))
    pos, axes, box, btype = _d(pos), _d(axes), _d(box), _i(btype)
    f, t = np.zeros_like(pos), np.zeros_like(pos)
    fn = lib().oxo_meta_coordination
    fn.restype = C.c_double
    val = fn(C.byref(c), pos.shape[0], _p(pos), _p(axes), _p(btype), _p(box), _p(f), _p(t))
    return dict(coordination=val, force=f, torque_lab=t)
