"""ctypes binding of the C ABI (include/oxdna_b200.h -> liboxdna_b200.so).

Nothing here computes: every call forwards host buffers to the CUDA library.  If the library cannot be loaded, or no
CUDA device is present, calls raise -- there is deliberately no CPU fallback in the product path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "liboxdna_b200.so")

PRECISION_FLOAT, PRECISION_MIXED = 0, 1
THERMOSTAT_NONE, THERMOSTAT_BROWNIAN, THERMOSTAT_LANGEVIN, THERMOSTAT_BUSSI = 0, 1, 2, 3
EXT_STRING, EXT_TRAP, EXT_MUTUAL_TRAP, EXT_LOWDIM_TRAP, EXT_REPULSION_PLANE, EXT_ATTRACTION_PLANE, EXT_SPHERE, EXT_LJ_WALL = range(8)
NTERMS = 8
NF4 = 13


class F1(C.Structure):
    _fields_ = [(n, C.c_float) for n in "a rc r0 blow bhigh rlow rhigh rclow rchigh".split()]


class F2(C.Structure):
    _fields_ = [(n, C.c_float) for n in "k rc r0 blow rlow rclow bhigh rhigh rchigh".split()]


class F4(C.Structure):
    _fields_ = [(n, C.c_float) for n in "a b t0 ts tc".split()]


class F5(C.Structure):
    _fields_ = [(n, C.c_float) for n in "a b xc xs".split()]


class Excl(C.Structure):
    _fields_ = [(n, C.c_float) for n in "sigma2 rstar2 b rc rc2".split()]


class DNA2Params(C.Structure):
    _fields_ = (
        [(n, C.c_float) for n in "back_a1 back_a2 stack_a1 base_a1 backref_a1 fene_eps fene_r0 fene_delta fene_delta2".split()]
        + [("use_mbf", C.c_int)]
        + [(n, C.c_float) for n in "mbf_xmax mbf_fmax mbf_finf mbf_e0 excl_eps".split()]
        + [("excl", Excl * 4), ("hb", F1), ("stck", F1)]
        + [(n, C.c_float * 25) for n in "hb_eps hb_shift stck_eps stck_shift".split()]
        + [("crst", F2), ("cxst", F2), ("f4", F4 * NF4), ("f4_cmin", C.c_float * NF4), ("f4_cmax", C.c_float * NF4), ("cxst_t1_sa", C.c_float), ("cxst_t1_sb", C.c_float), ("phi1", F5), ("phi2", F5)]
        + [(n, C.c_float) for n in "dh_minus_kappa dh_prefactor dh_rhigh dh_rc dh_b".split()]
        + [("dh_half_charged_ends", C.c_int), ("hb_multiplier", C.c_float), ("rcut", C.c_float), ("rcut_near", C.c_float), ("v1", C.c_int), ("phi3", F5)]
    )


NRF4 = 19


class RNA2Params(C.Structure):
    _fields_ = (
        [(n, C.c_float) for n in "back_a1 back_a2 back_a3 stack_a1 base_a1 stack3_a1 stack3_a2 stack5_a1 stack5_a2".split()]
        + [("p3", C.c_float * 3), ("p5", C.c_float * 3)]
        + [(n, C.c_float) for n in "fene_eps fene_r0 fene_delta fene_delta2".split()]
        + [("use_mbf", C.c_int)]
        + [(n, C.c_float) for n in "mbf_xmax mbf_fmax mbf_finf mbf_e0 excl_eps".split()]
        + [("excl", Excl * 4), ("hb", F1), ("stck", F1)]
        + [(n, C.c_float * 25) for n in "hb_eps hb_shift stck_eps stck_shift".split()]
        + [("crst", F2), ("cxst", F2), ("crst_kfac", C.c_float * 25), ("f4", F4 * NRF4), ("f4_cmin", C.c_float * NRF4), ("f4_cmax", C.c_float * NRF4)]
        + [(n, F5) for n in "phi1 phi2 phi3 phi4".split()]
        + [(n, C.c_float) for n in "dh_minus_kappa dh_prefactor dh_rhigh dh_rc dh_b".split()]
        + [("dh_half_charged_ends", C.c_int), ("average", C.c_int), ("mismatch_repulsion", C.c_int), ("mis_eps", C.c_float),
           ("mis_shift", C.c_float), ("hb_multiplier", C.c_float), ("rcut", C.c_float), ("rcut_near", C.c_float)]
    )


class ReplicaConsts(C.Structure):
    """oxb_replica_consts: the temperature-dependent constants of one replica of a batch"""
    _fields_ = ([("stck_eps", C.c_float * 25), ("stck_shift", C.c_float * 25)]
                + [(n, C.c_float) for n in "dh_minus_kappa dh_prefactor dh_rhigh dh_rc dh_b rcut2 th_a th_b th_c th_d".split()])


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


class ForceViews(C.Structure):
    """oxb_force_views (include/oxdna_b200.h): device pointers in the reference's layouts, handed to a plugged-in force pass"""
    _fields_ = [("N", C.c_int), ("stride", C.c_int), ("poss", C.c_void_p), ("orientations", C.c_void_p), ("matrix_neighs", C.c_void_p),
                ("number_neighs", C.c_void_p), ("bonds", C.c_void_p), ("forces", C.c_void_p), ("torques", C.c_void_p), ("box", C.c_double * 3),
                ("step", C.c_longlong), ("stream", C.c_void_p)]


class DNA3Scalars(C.Structure):
    """oxb_dna3_scalars (include/oxdna_b200.h): 29 doubles, the order of the block oracle/ref_harness.cpp:oxref_dna3_tables writes"""
    _fields_ = ([(n, C.c_double) for n in ("fene_eps use_mbf mbf_fmax mbf_finf hb_multiplier dh_rc dh_rhigh dh_prefactor dh_b dh_minus_kappa "
                                           "dh_half_charged_ends rcut").split()]
                + [("cxst_t1", C.c_double * 5), ("cxst_t4", C.c_double * 5), ("cxst_t5", C.c_double * 5), ("cxst_t1_sa", C.c_double), ("cxst_t1_sb", C.c_double)])


DNA3_NTAB, DNA3_TSIZE = 215, 900


def dna3_scalars(block):
    """from the flat block of 29 doubles (fixtures, reference harness)"""
    b = np.ascontiguousarray(block, dtype=np.float64)
    assert b.size == C.sizeof(DNA3Scalars) // 8
    return DNA3Scalars.from_buffer_copy(b.tobytes())


FORCE_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(ForceViews))

EXPORTED = """oxb_sizeof oxb_dna2_params_init oxb_dna1_params_init oxb_dna2_params_seqdep oxb_rna2_params_init oxb_rna2_params_seqdep oxb_set_model_rna2 oxb_set_model_dna3 oxb_create oxb_destroy oxb_last_error oxb_set_stream oxb_set_box
oxb_set_topology oxb_set_model_dna2 oxb_set_lists oxb_set_dt oxb_set_thermostat oxb_set_ext_forces oxb_set_ext_index_pool oxb_set_ext_grid_pool oxb_set_state oxb_get_state oxb_write_conf oxb_write_conf_binary
oxb_set_step oxb_get_step oxb_sort oxb_update_lists oxb_compute_forces oxb_first_step oxb_second_step oxb_thermostat oxb_run
oxb_synchronize oxb_get_forces oxb_energy oxb_barostat_move oxb_barostat_trial oxb_barostat_accept oxb_barostat_reject oxb_get_box oxb_set_host_wait oxb_fix_diffusion oxb_energy_split oxb_get_pairs oxb_get_stats oxb_device_views oxb_launch_count oxb_time_kernel oxb_set_profile oxb_get_profile
oxb_set_replicas oxb_set_replica_consts oxb_replica_energies oxb_replica_consts_dna2 oxb_replica_consts_rna2 oxb_set_force_callback""".split()

_lib = None


def lib():
    """Loads the CUDA library; raises if it is missing (build it with oxdna_b200/build.py or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} not built: oxdna_b200 has no fallback path, run `python oxdna_b200/build.py`")
        L = C.CDLL(SO_PATH)
        L.oxb_last_error.restype = C.c_char_p
        L.oxb_get_step.restype = C.c_longlong
        L.oxb_launch_count.restype = C.c_longlong
        L.oxb_last_error.argtypes = [C.c_void_p]
        L.oxb_destroy.argtypes = [C.c_void_p]
        L.oxb_destroy.restype = None
        L.oxb_replica_consts_dna2.restype = None
        L.oxb_replica_consts_rna2.restype = None
        for which, mirror in enumerate((DNA2Params, RNA2Params, ExtForce, ReplicaConsts)):
            if L.oxb_sizeof(which) != C.sizeof(mirror):
                raise RuntimeError(f"ABI mismatch: {mirror.__name__} is {C.sizeof(mirror)} bytes here, {L.oxb_sizeof(which)} in {SO_PATH}")
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class OxbError(RuntimeError):
    pass


def dna2_params(T, salt=0.5, dh_half_charged_ends=True, max_backbone_force=None, max_backbone_force_far=0.04):
    """oxb_dna2_params_init.  Returns (params, rcut)."""
    P = DNA2Params()
    rc = C.c_double()
    mbf = max_backbone_force is not None
    r = lib().oxb_dna2_params_init(C.byref(P), C.c_double(T), C.c_double(salt), int(dh_half_charged_ends), int(mbf),
                                   C.c_double(max_backbone_force if mbf else 0.0), C.c_double(max_backbone_force_far), C.byref(rc))
    if r != 0:
        raise OxbError("oxb_dna2_params_init failed")
    return P, rc.value


def dna1_params(T, grooving=False, max_backbone_force=None, max_backbone_force_far=0.04):
    """oxb_dna1_params_init (interaction_type = DNA / DNA_nomesh).  Returns (params, rcut)."""
    P = DNA2Params()
    rc = C.c_double()
    mbf = max_backbone_force is not None
    r = lib().oxb_dna1_params_init(C.byref(P), C.c_double(T), int(grooving), int(mbf), C.c_double(max_backbone_force if mbf else 0.0),
                                   C.c_double(max_backbone_force_far), C.byref(rc))
    if r != 0:
        raise OxbError("oxb_dna1_params_init failed")
    return P, rc.value


def dna2_params_seqdep(P, T, stck16, stck_fact_eps, hb_AT, hb_GC):
    """oxb_dna2_params_seqdep: STCK_X_Y table (A, G, C, T order), STCK_FACT_EPS, HYDR_A_T, HYDR_C_G of oxDNA2_sequence_dependent_parameters.txt."""
    s = _d(np.asarray(stck16).reshape(16))
    if lib().oxb_dna2_params_seqdep(C.byref(P), C.c_double(T), _p(s), C.c_double(stck_fact_eps), C.c_double(hb_AT), C.c_double(hb_GC)) != 0:
        raise OxbError("oxb_dna2_params_seqdep failed")
    return P


def rna2_params(T, salt=1.0, dh_half_charged_ends=True, max_backbone_force=None, max_backbone_force_far=0.04, mismatch_repulsion=False,
                mismatch_repulsion_strength=1.0):
    """oxb_rna2_params_init.  Returns (params, rcut)."""
    P = RNA2Params()
    rc = C.c_double()
    mbf = max_backbone_force is not None
    r = lib().oxb_rna2_params_init(C.byref(P), C.c_double(T), C.c_double(salt), int(dh_half_charged_ends), int(mbf),
                                   C.c_double(max_backbone_force if mbf else 0.0), C.c_double(max_backbone_force_far),
                                   int(mismatch_repulsion), C.c_double(mismatch_repulsion_strength), C.byref(rc))
    if r != 0:
        raise OxbError("oxb_rna2_params_init failed")
    return P, rc.value


def replica_consts(P, th=(0.0, 0.0, 0.0, 0.0)):
    """oxb_replica_consts_dna2 / _rna2: the temperature-dependent subset of a parameter block + thermostat constants a..d"""
    out = ReplicaConsts()
    fn = lib().oxb_replica_consts_rna2 if isinstance(P, RNA2Params) else lib().oxb_replica_consts_dna2
    fn(C.byref(P), C.c_double(th[0]), C.c_double(th[1]), C.c_double(th[2]), C.c_double(th[3]), C.byref(out))
    return out


def rna2_params_seqdep(P, T, stck16, st_t_dep, cross16, hb_AT, hb_GC, hb_GT):
    s, x = _d(np.asarray(stck16).reshape(16)), _d(np.asarray(cross16).reshape(16))
    if lib().oxb_rna2_params_seqdep(C.byref(P), C.c_double(T), _p(s), C.c_double(st_t_dep), _p(x), C.c_double(hb_AT), C.c_double(hb_GC),
                                    C.c_double(hb_GT)) != 0:
        raise OxbError("oxb_rna2_params_seqdep failed")
    return P


class Context:
    """One simulated system on one GPU (opaque oxb_ctx*)."""

    def __init__(self, N, device=0, precision=PRECISION_MIXED):
        self.N = int(N)
        self._h = C.c_void_p()
        self._L = lib()
        rc = self._L.oxb_create(C.byref(self._h), int(device), self.N, int(precision))
        if rc != 0:
            msg = self._L.oxb_last_error(self._h).decode() if self._h else "oxb_create failed"
            raise OxbError(msg)

    def close(self):
        if self._h:
            self._L.oxb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise OxbError(self._L.oxb_last_error(self._h).decode() + f" (code {rc})")

    # ---- configuration
    def set_box(self, box):
        b = _d(box)
        self._ck(self._L.oxb_set_box(self._h, _p(b)))

    def set_topology(self, btype, n3, n5, strand=None):
        a, b, c, d = _i(btype), _i(n3), _i(n5), _i(strand)
        self._ck(self._L.oxb_set_topology(self._h, _p(a), _p(b), _p(c), _p(d)))

    def set_model_dna2(self, P, rcut):
        self._ck(self._L.oxb_set_model_dna2(self._h, C.byref(P), C.c_double(rcut)))

    def set_model_rna2(self, P, rcut):
        self._ck(self._L.oxb_set_model_rna2(self._h, C.byref(P), C.c_double(rcut)))

    def set_model_dna3(self, tables, scalars):
        """tables: (215, 900) doubles in the order of include/oxdna_b200.h; scalars: DNA3Scalars or the flat block of 29 doubles"""
        t = np.ascontiguousarray(tables, dtype=np.float64)
        if t.size != DNA3_NTAB * DNA3_TSIZE:
            raise ValueError(f"oxDNA3 needs {DNA3_NTAB} tables of {DNA3_TSIZE} entries")
        S = scalars if isinstance(scalars, DNA3Scalars) else dna3_scalars(scalars)
        self._ck(self._L.oxb_set_model_dna3(self._h, _p(t), C.byref(S)))

    def set_replicas(self, n):
        self._ck(self._L.oxb_set_replicas(self._h, int(n)))
        self.n_replicas = int(n)

    def set_replica_consts(self, rows):
        arr = (ReplicaConsts * len(rows))(*rows)
        self._ck(self._L.oxb_set_replica_consts(self._h, len(rows), arr))

    def replica_energies(self):
        out = np.zeros(getattr(self, "n_replicas", 1))
        self._ck(self._L.oxb_replica_energies(self._h, _p(out)))
        return out

    def set_lists(self, verlet_skin=0.05, use_edge=False, sort_every=0, max_density_multiplier=3.0):
        self._ck(self._L.oxb_set_lists(self._h, C.c_double(verlet_skin), int(use_edge), int(sort_every), C.c_double(max_density_multiplier)))

    def set_dt(self, dt):
        self._ck(self._L.oxb_set_dt(self._h, C.c_double(dt)))

    def set_thermostat(self, kind, every=1, a=0.0, b=0.0, c=0.0, d=0.0, seed=0):
        self._ck(self._L.oxb_set_thermostat(self._h, int(kind), int(every), C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(d),
                                            C.c_ulonglong(seed)))

    def set_ext_forces(self, forces):
        n = len(forces)
        arr = (ExtForce * max(n, 1))()
        pool, grid = [], []
        for k, f in enumerate(forces):
            fill_ext_entry(arr[k], f, pool, grid)
        self._ck(self._L.oxb_set_ext_forces(self._h, 0, arr))  # COM entries refer to the pools: drop them before replacing those
        self._ck(self._L.oxb_set_ext_index_pool(self._h, len(pool), (C.c_int * max(len(pool), 1))(*pool)))
        self._ck(self._L.oxb_set_ext_grid_pool(self._h, len(grid), (C.c_double * max(len(grid), 1))(*grid)))
        self._ck(self._L.oxb_set_ext_forces(self._h, n, arr))

    # ---- state
    def set_state(self, pos, a1, a3, vel=None, L=None):
        a, b, c, d, e = _d(pos), _d(a1), _d(a3), _d(vel), _d(L)
        self._ck(self._L.oxb_set_state(self._h, _p(a), _p(b), _p(c), _p(d), _p(e)))

    def get_state(self, out=None):
        """device state -> dict(pos, a1, a3, vel, L) of (N, 3) float64 arrays; `out` may supply the destination arrays (e.g. pinned
        host memory, which the device copies straight into)"""
        keys = ("pos", "a1", "a3", "vel", "L")
        if out is None:
            out = {k: np.empty((self.N, 3)) for k in keys}
        for k in keys:
            a = out[k]
            if a.dtype != np.float64 or a.shape != (self.N, 3) or not a.flags.c_contiguous:
                raise ValueError(f"get_state: out['{k}'] must be a C-contiguous float64 array of shape ({self.N}, 3)")
        self._ck(self._L.oxb_get_state(self._h, *[_p(out[k]) for k in keys]))
        return out

    def write_conf(self, path, append=False, print_momenta=True):
        """one frame in the reference's configuration format, written from the device state"""
        self._ck(self._L.oxb_write_conf(self._h, str(path).encode(), int(append), int(print_momenta)))

    def write_conf_binary(self, path, append=False, rng_state=None, pos_shift=None):
        """one frame in the reference's binary configuration format"""
        seed = None if rng_state is None else (C.c_ushort * 3)(*[int(x) for x in rng_state])
        sh = None if pos_shift is None else np.ascontiguousarray(pos_shift, dtype=np.int32)
        self._ck(self._L.oxb_write_conf_binary(self._h, str(path).encode(), int(append), seed, _p(sh)))

    def set_step(self, s):
        self._ck(self._L.oxb_set_step(self._h, C.c_longlong(s)))

    @property
    def step(self):
        return self._L.oxb_get_step(self._h)

    # ---- operators
    def sort(self):
        self._ck(self._L.oxb_sort(self._h))

    def update_lists(self):
        self._ck(self._L.oxb_update_lists(self._h))

    def compute_forces(self):
        self._ck(self._L.oxb_compute_forces(self._h))

    def first_step(self):
        self._ck(self._L.oxb_first_step(self._h))

    def second_step(self):
        self._ck(self._L.oxb_second_step(self._h))

    def thermostat(self):
        self._ck(self._L.oxb_thermostat(self._h))

    def run(self, n):
        self._ck(self._L.oxb_run(self._h, C.c_longlong(n)))

    def synchronize(self):
        self._ck(self._L.oxb_synchronize(self._h))

    # ---- read-backs
    def get_forces(self):
        N = self.N
        f, tb, tl, e, hb = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N), np.zeros(N)
        self._ck(self._L.oxb_get_forces(self._h, _p(f), _p(tb), _p(tl), _p(e), _p(hb)))
        return dict(force=f, torque_body=tb, torque_lab=tl, energy=e, hb_energy=hb, U=0.5 * e.sum())

    def energy(self):
        U, K = C.c_double(), C.c_double()
        self._ck(self._L.oxb_energy(self._h, C.byref(U), C.byref(K)))
        return U.value, K.value

    def barostat_move(self, new_box, molecular, P, T, u):
        """one MC volume move (MD_CUDABackend::_apply_barostat); returns (accepted, dE)"""
        acc, dE = C.c_int(), C.c_double()
        b = _d(new_box)
        self._ck(self._L.oxb_barostat_move(self._h, _p(b), int(bool(molecular)), C.c_double(P), C.c_double(T), C.c_double(u), C.byref(acc), C.byref(dE)))
        return bool(acc.value), dE.value

    def barostat_trial(self, new_box, molecular):
        b = _d(new_box)
        self._ck(self._L.oxb_barostat_trial(self._h, _p(b), int(bool(molecular))))

    def barostat_accept(self):
        self._ck(self._L.oxb_barostat_accept(self._h))

    def barostat_reject(self):
        self._ck(self._L.oxb_barostat_reject(self._h))

    def fix_diffusion(self):
        """strands back into the box by whole box sides; returns the (N, 3) int shifts floor(com / L)"""
        sh = np.zeros((self.N, 3), dtype=np.int32)
        self._ck(self._L.oxb_fix_diffusion(self._h, _p(sh)))
        return sh

    def set_host_wait(self, blocking):
        self._ck(self._L.oxb_set_host_wait(self._h, int(bool(blocking))))

    def get_box(self):
        b = np.zeros(3)
        self._ck(self._L.oxb_get_box(self._h, _p(b)))
        return b

    def energy_split(self):
        """per-term potential energies (FENE, BEXC, STCK, NEXC, HB, CRSTCK, CXSTCK, DH), summed on the device"""
        out = np.zeros(NTERMS)
        self._ck(self._L.oxb_energy_split(self._h, _p(out)))
        return out

    def get_pairs(self):
        n = C.c_longlong()
        self._ck(self._L.oxb_get_pairs(self._h, None, C.c_longlong(0), C.byref(n)))
        out = np.zeros((max(n.value, 1), 2), dtype=np.int32)
        self._ck(self._L.oxb_get_pairs(self._h, _p(out), C.c_longlong(n.value), C.byref(n)))
        return out[: n.value]

    def device_views(self):
        """oxb_device_views: raw device pointers (ints) of the reference-layout views, for zero-copy consumers (plugin seam)"""
        p = [C.c_void_p() for _ in range(6)]
        self._ck(self._L.oxb_device_views(self._h, *[C.byref(x) for x in p]))
        return dict(zip(("poss", "orientations", "matrix_neighs", "number_neighs", "edge_list", "n_edges"), [x.value for x in p]))

    def set_force_callback(self, fn, rcut):
        """oxb_set_force_callback: fn(views: ForceViews) -> int enqueues a third-party force pass on views.stream (None removes it)"""
        if fn is None:
            self._force_cb = None
            self._ck(self._L.oxb_set_force_callback(self._h, None, None, C.c_double(0.0)))
            return

        def tramp(_user, vp):
            try:
                return int(fn(vp.contents) or 0)
            except Exception:  # an exception cannot cross the C frame
                import traceback
                traceback.print_exc()
                return 1
        self._force_cb = FORCE_CALLBACK(tramp)  # keep the thunk alive
        self._ck(self._L.oxb_set_force_callback(self._h, self._force_cb, None, C.c_double(rcut)))

    def stats(self):
        a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_int(), C.c_int()
        self._ck(self._L.oxb_get_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(n_list_updates=a.value, n_sorts=b.value, max_neigh=c.value, error_flags=d.value)

    def launch_count(self):
        return self._L.oxb_launch_count(self._h)

    PROF_PHASES = ("other", "force", "integrate", "wait", "sort", "build", "gap", "permute", "edges")

    def set_profile(self, enable=True):
        self._ck(self._L.oxb_set_profile(self._h, int(bool(enable))))

    def get_profile(self):
        """{phase: (milliseconds, entries)} accumulated inside run() since set_profile(True)"""
        ms, n = (C.c_double * len(self.PROF_PHASES))(), (C.c_longlong * len(self.PROF_PHASES))()
        self._ck(self._L.oxb_get_profile(self._h, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.PROF_PHASES)}

    def time_kernel(self, which, reps=10):
        ms = C.c_float()
        self._ck(self._L.oxb_time_kernel(self._h, int(which), int(reps), C.byref(ms)))
        return ms.value
