"""Python mirror of the reference's input-file driven set-up for the GPU MD path (`backend = CUDA`).

`Simulation(inp, topology, conf)` accepts the reference's input keys (docs/source/input.md; the CUDA ones are read in
src/CUDA/Backends/CUDABaseBackend.cu:110-141, MD_CUDABackend.cu:621-672, CUDAThermostatFactory.cu:18-45) and drives
the C ABI.  It is the host-side convenience used by tests, bench.py and the REMD driver; the C++ host classes in
oxdna_b200/host mirror the same interface for linking against the reference's SimManager.
"""
import types

import numpy as np

from . import capi


def parse_temperature(raw):
    """src/Utilities/Utils.cpp:316-346: '300K', '27C' or a number in simulation units."""
    if isinstance(raw, (int, float)):
        return float(raw)
    s = str(raw).strip()
    if s[-1] in "kK":
        return float(s[:-1]) * 0.1 / 300.0
    if s[-1] in "cC":
        return (float(s[:-1]) + 273.15) * 0.1 / 300.0
    return float(s)


def _bool(v):
    if isinstance(v, str):
        return v.strip().lower() in ("1", "true", "yes", "on")
    return bool(v)


def brownian_params(T, dt, newtonian_steps, pt=0.0, diff_coeff=0.0):
    """BrownianThermostat::init (src/Backends/Thermostats/BrownianThermostat.cpp:43-54); pt/diff_coeff/dt are floats there."""
    dt, D, p = float(np.float32(dt)), float(np.float32(diff_coeff)), float(np.float32(pt))
    if p == 0.0:
        p = (2 * T * newtonian_steps * dt) / (T * newtonian_steps * dt + 2 * D)
    if p > 1.0:
        raise ValueError(f"pt ({p}) must be smaller than 1")
    D = T * newtonian_steps * dt * (1.0 / p - 0.5)
    pr = (2 * T * newtonian_steps * dt) / (T * newtonian_steps * dt + 2 * 3 * D)
    return p, pr, np.sqrt(T)


def langevin_params(T, dt, gamma_trans=0.0, diff_coeff=0.0):
    """LangevinThermostat::init (src/Backends/Thermostats/LangevinThermostat.cpp:57-76)."""
    g, D = float(np.float32(gamma_trans)), float(np.float32(diff_coeff))
    if g != 0.0 and D != 0.0:
        raise ValueError("Cannot specify both gamma_trans and diff_coeff. Remove one of them from the input file")
    if g == 0.0 and D == 0.0:
        raise ValueError("Must specify one of gamma_trans and diff_coeff for the Langevin thermostat to work")
    if D == 0.0:
        D = T / g
    else:
        g = T / D
    gr = T / (3.0 * D)
    return g, gr, np.sqrt(2.0 * g * T / dt), np.sqrt(2.0 * gr * T / dt)


def read_seq_dep(path_or_dict):
    """`seq_dep_file` (key = value lines, e.g. oxDNA2_sequence_dependent_parameters.txt / rna_sequence_dependent_parameters.txt),
    or an already parsed dict."""
    if isinstance(path_or_dict, dict):
        return path_or_dict
    if path_or_dict is None:
        raise ValueError("use_average_seq = false needs seq_dep_file")
    out = {}
    with open(path_or_dict) as f:
        for line in f:
            line = line.split("#")[0]
            if "=" in line:
                k, v = line.split("=", 1)
                out[k.strip()] = float(np.float32(v))  # the reference reads these with getInputFloat
    return out


class Simulation:
    def __init__(self, inp, topology, conf, device=0):
        """topology: dict(btype, n3, n5, strand); conf: dict(box, pos, a1, a3[, vel, L])."""
        self.inp = dict(inp)
        g = self.inp.get
        if str(g("backend", "CUDA")).upper() != "CUDA":
            raise ValueError("oxdna_b200 only implements backend = CUDA")
        itype = str(g("interaction_type", "DNA2"))
        if itype not in ("DNA2", "DNA2_nomesh", "DNA", "DNA_nomesh", "RNA2", "RNA", "DNA3", "DNA3_nomesh"):
            raise ValueError(f"interaction_type = {itype} is not available in this build (DNA, DNA2, DNA3, RNA, RNA2)")
        self.itype = itype
        prec = str(g("backend_precision", "mixed"))
        if prec not in ("mixed", "float"):
            raise ValueError(f"backend_precision = {prec} is not available (float and mixed are; float is served by the mixed kernels)")
        if prec == "double" and _bool(g("use_edge", 0)):
            raise ValueError("use_edge and double precision are not compatible")  # MD_CUDABackend.cu:625-632
        if "reload_from" in self.inp:
            # CUDABaseBackend.cu:137-140
            raise ValueError("The CUDA backend does not support checkpoints (reload_from)")
        self.use_edge = _bool(g("use_edge", 0))
        if str(g("CUDA_list", "verlet")) == "no" and self.use_edge:
            raise ValueError("'CUDA_list = no' and 'use_edge = true' are incompatible")  # CUDANoList.cu:20-26
        self.T = parse_temperature(g("T"))
        self.dt = float(g("dt", 0.003))
        self.N = N = len(topology["btype"])
        self.ctx = capi.Context(N, device=device, precision=capi.PRECISION_FLOAT if prec == "float" else capi.PRECISION_MIXED)
        c = self.ctx
        c.set_box(conf["box"])
        c.set_topology(topology["btype"], topology["n3"], topology["n5"], topology.get("strand"))
        self._set_model()
        c.set_lists(float(g("verlet_skin", 0.05)), self.use_edge, int(g("CUDA_sort_every", 0)), float(g("max_density_multiplier", 3.0)))
        c.set_dt(self.dt)
        self.seed = int(g("seed", 42))
        self._set_thermostat()
        ext = g("external_forces_list", None)
        if ext:
            c.set_ext_forces(ext)
        c.set_state(conf["pos"], conf["a1"], conf["a3"], conf.get("vel"), conf.get("L"))
        # MC barostat, MDBackend::get_settings (src/Backends/MDBackend.cpp:43-59)
        self.use_barostat = _bool(g("use_barostat", 0))
        if self.use_barostat:
            self.P, self.delta_L = float(g("P")), float(g("delta_L"))
            self.barostat_probability = float(g("barostat_probability"))
            self.barostat_isotropic = _bool(g("barostat_isotropic", 1))
            self.barostat_molecular = _bool(g("barostat_molecular", 0))
            self.barostat_attempts = self.barostat_accepted = 0
            self._rng = np.random.default_rng(self.seed + 7919)

    def _model_for(self, T):
        """(parameter block, rcut) of the configured interaction at temperature T (simulation units)"""
        g = self.inp.get
        mbf = g("max_backbone_force", None)
        mbf = None if mbf is None else float(mbf)
        average = _bool(g("use_average_seq", 1))
        sd = None if average else read_seq_dep(g("seq_dep_file"))
        if self.itype in ("RNA2", "RNA"):
            # RNA2Interaction::get_settings: salt defaults to 1.0 (src/Interactions/RNAInteraction2.cpp:34-37); interaction_type = RNA
            # (class RNAInteraction) is the same model without the Debye-Hueckel and mismatch terms: salt 0 switches them off
            v2 = self.itype == "RNA2"
            params, rcut = capi.rna2_params(T, float(g("salt_concentration", 1.0)) if v2 else 0.0, _bool(g("dh_half_charged_ends", 1)), mbf,
                                            float(g("max_backbone_force_far", 0.04)), v2 and _bool(g("mismatch_repulsion", 0)),
                                            float(g("mismatch_repulsion_strength", 1.0)))
            if sd is not None:
                B = "AGCT"
                hb = lambda a, b: sd.get(f"HYDR_{a}_{b}", sd.get(f"HYDR_{b}_{a}"))
                capi.rna2_params_seqdep(params, T, [sd[f"STCK_{a}_{b}"] for a in B for b in B], sd["ST_T_DEP"],
                                        [sd[f"CROSS_{a}_{b}"] for a in B for b in B], hb("A", "T"), hb("G", "C"), hb("G", "T"))
            return params, rcut
        if self.itype in ("DNA", "DNA_nomesh"):
            params, rcut = capi.dna1_params(T, _bool(g("major_minor_grooving", 0)), mbf, float(g("max_backbone_force_far", 0.04)))
        else:
            params, rcut = capi.dna2_params(T, float(g("salt_concentration", 0.5)), _bool(g("dh_half_charged_ends", 1)), mbf,
                                            float(g("max_backbone_force_far", 0.04)))
        if sd is not None:
            # DNAInteraction.cpp:329-375
            B = "AGCT"
            hb = lambda a, b: sd.get(f"HYDR_{a}_{b}", sd.get(f"HYDR_{b}_{a}"))
            capi.dna2_params_seqdep(params, T, [sd[f"STCK_{a}_{b}"] for a in B for b in B], sd["STCK_FACT_EPS"], hb("A", "T"), hb("G", "C"))
        return params, rcut

    def _set_model(self):
        if self.itype in ("DNA3", "DNA3_nomesh"):
            # oxDNA3: the tetramer-indexed tables are input (what DNA3Interaction::init leaves in the class and CUDADNA3Interaction::cuda_init
            # uploads, CUDADNA3Interaction.cu:46-150): keys dna3_tables (215 x 900) and dna3_scalars (29 doubles, capi.DNA3Scalars)
            tab, sc = self.inp.get("dna3_tables"), self.inp.get("dna3_scalars")
            if tab is None or sc is None:
                raise ValueError("interaction_type = DNA3 needs the parameter tables of DNA3Interaction (dna3_tables, dna3_scalars)")
            S = capi.dna3_scalars(sc) if not isinstance(sc, capi.DNA3Scalars) else sc
            self.rcut = float(S.rcut)
            self.ctx.set_model_dna3(tab, S)
            # what callers read from a parameter block (bench.pair_statistics): backbone-site offsets (src/model.h:16-17) and radii
            self.params = types.SimpleNamespace(scalars=S, back_a1=-0.34, back_a2=0.3408, dh_rc=float(S.dh_rc), rcut=self.rcut, rcut_near=self.rcut)
            return
        self.params, self.rcut = self._model_for(self.T)
        if self.itype in ("RNA2", "RNA"):
            self.ctx.set_model_rna2(self.params, self.rcut)
        else:
            self.ctx.set_model_dna2(self.params, self.rcut)

    def _thermostat_for(self, T):
        """(type, every, a, b, c, d) of oxb_set_thermostat for the configured thermostat at temperature T"""
        g = self.inp.get
        kind = str(g("thermostat", "no")).lower()
        if kind == "no":
            return (capi.THERMOSTAT_NONE, 1, 0.0, 0.0, 0.0, 0.0)
        if kind in ("john", "brownian"):  # synonyms, CUDAThermostatFactory.cu:23-28
            ns = int(g("newtonian_steps"))
            pt, pr, resc = brownian_params(T, self.dt, ns, float(g("pt", 0.0)), float(g("diff_coeff", 0.0)))
            return (capi.THERMOSTAT_BROWNIAN, ns, pt, pr, resc, 0.0)
        if kind == "langevin":
            gt, gr, rt, rr = langevin_params(T, self.dt, float(g("gamma_trans", 0.0)), float(g("diff_coeff", 0.0)))
            return (capi.THERMOSTAT_LANGEVIN, 1, gt, gr, rt, rr)
        if kind == "bussi":
            ns, tau = int(g("newtonian_steps")), int(g("bussi_tau"))
            return (capi.THERMOSTAT_BUSSI, ns, T, np.exp(-ns / float(tau)), 0.0, 0.0)
        raise ValueError(f"Invalid thermostat '{kind}'")

    def _set_thermostat(self):
        kind, every, a, b, c, d = self._thermostat_for(self.T)
        self.ctx.set_thermostat(kind, every, a, b, c, d, self.seed)

    def update_temperature(self, T, dna3_tables=None, dna3_scalars=None):
        """ConfigInfo::update_temperature -> interaction re-init + thermostat re-init.  oxDNA3: the tables depend on T
        (stacking strengths) and are input here -- hand in the ones DNA3Interaction::init derives for the new temperature."""
        if self.itype in ("DNA3", "DNA3_nomesh"):
            if dna3_tables is None or dna3_scalars is None:
                raise ValueError("interaction_type = DNA3: update_temperature needs the parameter tables for the new temperature")
            self.inp["dna3_tables"], self.inp["dna3_scalars"] = dna3_tables, dna3_scalars
        self.T = parse_temperature(T)
        self._set_model()
        self._set_thermostat()

    def run(self, steps):
        if not self.use_barostat:
            self.ctx.run(steps)
            return
        # MD_CUDABackend::sim_step (src/CUDA/Backends/MD_CUDABackend.cu:595-599): every step the barostat fires with probability
        # barostat_probability (MDBackend::_is_barostat_active); the steps in between go out as one fused run
        fire = np.flatnonzero(self._rng.random(steps) < self.barostat_probability)
        done = 0
        for s in fire:
            self.ctx.run(int(s) + 1 - done)
            done = int(s) + 1
            self.barostat_attempt()
        self.ctx.run(steps - done)

    def barostat_attempt(self):
        """MD_CUDABackend::_apply_barostat (src/CUDA/Backends/MD_CUDABackend.cu:451-516)"""
        box = self.ctx.get_box()
        if self.barostat_isotropic:
            new = box + self.delta_L * (self._rng.random() - 0.5)
        else:
            new = box + self.delta_L * (self._rng.random(3) - 0.5)
        ok, _ = self.ctx.barostat_move(new, self.barostat_molecular, self.P, self.T, self._rng.random())
        self.barostat_attempts += 1
        self.barostat_accepted += int(ok)
        return ok

    def system_energy(self):
        return self.ctx.energy()[0]

    def close(self):
        self.ctx.close()


class ReplicaBatch(Simulation):
    """R temperature replicas of one system in ONE GPU context (replica batching, include/oxdna_b200.h oxb_set_replicas): every kernel of
    the step is launched once for all replicas; the temperature-dependent constants (stacking strength, Debye-Hueckel, thermostat) live
    in a per-replica device table, so a temperature swap rewrites table rows and nothing else.  The reference runs one process and one
    GPU context per replica (examples/OXPY_REMD/remd.py:67-102) and re-initialises interaction + thermostat on every accepted swap.

    topology: ONE replica's topology; confs: one conf dict per replica (same box); temperatures: simulation units, one per replica;
    ladder_max: hottest temperature any replica may be given later (fixes the list radii; default max(temperatures))."""

    def __init__(self, inp, topology, confs, temperatures, device=0, ladder_max=None):
        R, n = len(confs), len(topology["btype"])
        if len(temperatures) != R:
            raise ValueError("one temperature per replica")
        self.n_replicas, self.n_per = R, n
        off = (np.arange(R) * n)[:, None]

        def rep_idx(a):  # neighbour indices: -1 stays -1
            a = np.asarray(a)[None, :]
            return np.where(a >= 0, a + off, a).reshape(-1)

        nstrand = int(np.max(topology.get("strand", np.zeros(n, dtype=int)))) + 1
        top = dict(btype=np.tile(np.asarray(topology["btype"]), R), n3=rep_idx(topology["n3"]), n5=rep_idx(topology["n5"]),
                   strand=(np.asarray(topology.get("strand", np.zeros(n, dtype=int)))[None, :] + (np.arange(R) * nstrand)[:, None]).reshape(-1))
        cat = lambda k: None if confs[0].get(k) is None else np.concatenate([np.asarray(c[k], dtype=np.float64) for c in confs])
        conf = dict(box=confs[0]["box"], pos=cat("pos"), a1=cat("a1"), a3=cat("a3"), vel=cat("vel"), L=cat("L"))
        inp = dict(inp)
        if str(inp.get("thermostat", "no")).lower() == "bussi":
            raise ValueError("replica batching is not available with the Bussi thermostat")
        ext = inp.get("external_forces_list")
        if ext:
            full = []
            for r in range(R):
                for e in ext:
                    e = dict(e)
                    for key in ("particle", "ref_particle"):
                        if isinstance(e.get(key), (int, np.integer)) and e[key] >= 0:
                            e[key] = int(e[key]) + r * n
                        elif key in e and not isinstance(e[key], (int, np.integer)):
                            raise ValueError("replica batching: external forces must name single particles")
                    full.append(e)
            inp["external_forces_list"] = full
        self.temps = np.asarray(temperatures, dtype=np.float64)
        inp["T"] = float(max(float(np.max(self.temps)), ladder_max or 0.0))  # list radii: the hottest Hamiltonian of the ladder
        super().__init__(inp, top, conf, device=device)
        self.ctx.set_replicas(R)
        self._rows = {}
        self._table = None
        self.set_temperatures(self.temps)

    def _row(self, T):
        key = float(T)
        if key not in self._rows:
            P, _ = self._model_for(key)
            self._rows[key] = capi.replica_consts(P, self._thermostat_for(key)[2:])
        return self._rows[key]

    def set_temperatures(self, T, first=False):
        T = np.asarray(T, dtype=np.float64)
        if self._table is None or not np.array_equal(T, self._table):
            self.ctx.set_replica_consts([self._row(t) for t in T])
            self._table = T.copy()
        self.temps = T.copy()

    def energies(self):
        return self.ctx.replica_energies()

    def energies_at(self, T, mask=None):
        """potential energy of every replica under the Hamiltonian of temperature T[r] (one force pass for the whole batch); the table
        stays at T until set_temperatures is called"""
        T = np.asarray(T, dtype=np.float64)
        if not np.array_equal(T, self._table):
            self.ctx.set_replica_consts([self._row(t) for t in T])
            self._table = T.copy()
        return self.ctx.replica_energies()

    def system_energy(self):
        return float(np.sum(self.energies()))

    def update_temperature(self, T):
        raise ValueError("a replica batch takes one temperature per replica: set_temperatures")

    def get_states(self):
        st = self.ctx.get_state()
        n = self.n_per
        return [{k: v[r * n:(r + 1) * n] for k, v in st.items()} for r in range(self.n_replicas)]


def make_batches(inp, topology, confs, temperatures, device=0, ladder_max=None, max_particles=4000000):
    """Splits the local replicas into as few ReplicaBatch contexts as the 22-bit particle index allows (64 x 81,920 nt = 2 batches)."""
    R, n = len(confs), len(topology["btype"])
    per = max(1, min(R, max_particles // n))
    n_batches = (R + per - 1) // per
    per = (R + n_batches - 1) // n_batches
    out = []
    for b in range(0, R, per):
        out.append(ReplicaBatch(inp, topology, confs[b:b + per], temperatures[b:b + per], device=device, ladder_max=ladder_max))
    return out
