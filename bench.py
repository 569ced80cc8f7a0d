#!/usr/bin/env python
"""Benchmark of the oxDNA GPU MD step (BASELINE.json metric: particle-steps/s).

  python bench.py --gpus 1 --steps K --warmup W              our CUDA path, N = 1: config C4 (1M-nt duplex lattice, DH, 50,000 mutual traps) as
                                                              the headline; C2, C3 and C5-on-one-GPU under "extras"
  torchrun ... bench.py --gpus N ...                          N > 1: config C5 -- 64 temperature replicas of C2 (290-350 K), 64/N per GPU as
                                                              replica batches, replica exchange over NCCL every bench step
  python bench.py --impl reference ...                        the reference's own CPU implementation of the same workload on the host cores

One bench "step" = one block of `md_steps_per_step` MD steps (the unit the reference's OxpyManager.run(steps) /
REMD pt_move_every works in); value = N_particles * md_steps / device time.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STR, SALT, DT = "300K", 0.5, 0.003
_REAL_STDOUT = sys.stdout
FLOP_FAR, FLOP_DH, FLOP_CONTACT, FLOP_BONDED = 170.0, 60.0, 1500.0, 900.0  # SURVEY 8(d) per-pair figures
# ncu dram__bytes_read.sum + dram__bytes_write.sum of one force pass, summed over its kernels (profiles/summary_r02c.txt: one
# `ncu --set full` capture per workload, cold cache: the compulsory traffic of one pass)
#   C2: Debye-Hueckel, bonded, near edges incl. the folded HB / cross-stacking / coaxial / FP64 excluded-volume tails
#   C4: Debye-Hueckel, bonded, near edges incl. coaxial / FP64 excluded-volume tails, HB + cross stacking, mutual traps
NCU_FORCE_PASS_DRAM_BYTES_C2 = int((7.80 + 7.29 + 9.08) * 1e6)
NCU_FORCE_PASS_DRAM_BYTES_C4 = int((99.73 + 94.58 + 88.36 + 127.72 + 18.99) * 1e6)

def workload(name):
    from oxdna_b200 import lattice
    if name == "c2":
        sysm = lattice.duplex_lattice(2048, bp=20, spacing=10.0, seed=12345)  # 13^3 sites, L = 130
        desc = "C2: oxDNA2 synthetic lattice of 2,048 x 20-bp duplexes (81,920 nt), L=130, salt 0.5, T=300K"
    elif name == "c4":
        sysm = lattice.duplex_lattice(25000, bp=20, spacing=10.0, seed=12345, sites_per_side=30)
        desc = "C4: oxDNA2 1M-nt lattice (25,000 x 20-bp), L=300, salt 0.5, 50,000 mutual traps"
    elif name == "c3":
        sysm = lattice.rna_duplex_lattice(2048, bp=16, spacing=10.0, seed=12345)  # 13^3 sites, L = 130
        desc = "C3: oxRNA2 synthetic lattice of 2,048 x 16-bp A-form duplexes (65,536 nt), sequence-dependent parameters, L=130, salt 0.5, T=300K"
    elif name == "c2_dna3":
        sysm = lattice.duplex_lattice(2048, bp=20, spacing=10.0, seed=12345)
        desc = "C2 geometry under oxDNA3 (interaction_type = DNA3, average-sequence tetramer tables): 2,048 x 20-bp duplexes (81,920 nt), L=130, salt 0.5, T=300K"
    elif name == "c4_dna3":
        sysm = lattice.duplex_lattice(25000, bp=20, spacing=10.0, seed=12345, sites_per_side=30)
        desc = "C4 geometry under oxDNA3 (interaction_type = DNA3, average-sequence tetramer tables): 1M nt (25,000 x 20-bp), L=300, salt 0.5, T=300K, no external forces"
    elif name == "small":
        sysm = lattice.duplex_lattice(64, bp=20, spacing=10.0, seed=12345)
        desc = "small: 64 x 20-bp duplexes (2,560 nt)"
    else:
        raise ValueError(name)
    return sysm, desc


def model_keys(name, tmpdir=None):
    """interaction keys of the workload; with tmpdir the sequence-dependent table is written to a file (reference binaries)"""
    if name in ("c2_dna3", "c4_dna3"):
        # oxDNA3 with use_average_seq = 1 (the reference binaries fill their tables without a parameter file); our side takes the tables the
        # reference CPU class derives for T = 300 K, salt 0.5 from the committed fixture (oracle/make_golden.py dna3)
        if tmpdir is not None:
            return dict(interaction_type="DNA3")
        g = np.load(os.path.join(ROOT, "tests", "golden", "dna3_tables_avg_300K_salt05.npz"))
        return dict(interaction_type="DNA3", dna3_tables=g["dna3_tables"], dna3_scalars=g["dna3_scalars"])
    if name != "c3":
        return dict(interaction_type="DNA2")
    from oxdna_b200 import seqdep
    sd = seqdep.RNA_SEQ_DEP if tmpdir is None else seqdep.write_file(os.path.join(tmpdir, "rna_seq_dep.txt"), seqdep.RNA_SEQ_DEP)
    return dict(interaction_type="RNA2", use_average_seq=0, seq_dep_file=sd)


def base_input(args, T=T_STR):
    d = dict(backend="CUDA", backend_precision="mixed", T=T, salt_concentration=SALT, dt=DT,
             verlet_skin=0.05, thermostat="brownian", newtonian_steps=103, diff_coeff=2.5, CUDA_list="verlet",
             CUDA_sort_every=args.sort_every, use_edge=args.use_edge, seed=42)
    d.update(model_keys(args.workload))
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm))


def pair_statistics(sim, sysm):
    """Unique listed pairs / DH-range pairs / near pairs per particle of the current configuration (for the FLOP model)."""
    st = sim.ctx.get_state()
    pairs = sim.ctx.get_pairs()
    N = sim.N
    from oxdna_b200 import io as oio
    ax = oio.orthonormal_axes(st["a1"], st["a3"])
    P = sim.params
    back = st["pos"] + ax[:, 0:3] * float(P.back_a1) + ax[:, 3:6] * float(P.back_a2) + ax[:, 6:9] * float(getattr(P, "back_a3", 0.0))
    box = sysm["box"]
    d = back[pairs[:, 1]] - back[pairs[:, 0]]
    d -= np.rint(d / box) * box
    rb = np.linalg.norm(d, axis=1)
    dc = st["pos"][pairs[:, 1]] - st["pos"][pairs[:, 0]]
    dc -= np.rint(dc / box) * box
    rc = np.linalg.norm(dc, axis=1)
    return dict(listed=len(pairs) / N, dh=float(np.sum(rb < float(sim.params.dh_rc))) / N, near=float(np.sum(rc < float(sim.params.rcut_near))) / N,
                contact=float(np.sum(rc < 1.0)) / N)


# ----------------------------------------------------------------------------------------------------------- CPU baseline
def write_case(sysm, T, d, state=None):
    """Topology + configuration files of the workload; `state` (a get_state() dict) replaces the ideal lattice."""
    from oxdna_b200 import io as oio, lattice
    from oxdna_b200.sim import parse_temperature
    top, conf = os.path.join(d, "bench.top"), os.path.join(d, "bench.dat")
    oio.write_topology(top, sysm["btype"], sysm["n3"], sysm["n5"], sysm["strand"])
    if state is None:
        v, L = lattice.maxwell_velocities(len(sysm["pos"]), parse_temperature(T), 5)
        oio.write_conf(conf, sysm["box"], sysm["pos"], sysm["a1"], sysm["a3"], v, L)
    else:
        oio.write_conf(conf, sysm["box"], state["pos"], state["a1"], state["a3"], state["vel"], state["L"])
    return top, conf


def _ref_worker(kind, top, conf, md_steps, warm, steps, barrier, q, keys=None):
    """One single-threaded CPU MD process (the reference is single-threaded by design)."""
    try:
        if kind == "reference":
            from oracle.refharness import Reference
            r = Reference(top, conf, salt_concentration=SALT, T=T_STR, thermostat="brownian", newtonian_steps=103,
                          diff_coeff=2.5, dt=DT, seed=42, **(keys or dict(interaction_type="DNA2")))
            stepper = r.step
        else:
            from oracle import oracle as O
            from oxdna_b200 import io as oio
            from oxdna_b200.sim import parse_temperature
            t, c = oio.read_topology(top), oio.read_conf(conf)
            if keys and keys.get("interaction_type") == "RNA2":
                from oxdna_b200.sim import read_seq_dep
                sd, B = read_seq_dep(keys["seq_dep_file"]), "AGCT"
                P = O.rna2_params(parse_temperature(T_STR), SALT, cpu_quirks=True)
                O.rna2_params_seqdep(P, [sd[f"STCK_{a}_{b}"] for a in B for b in B], sd["ST_T_DEP"], [sd[f"CROSS_{a}_{b}"] for a in B for b in B],
                                     sd["HYDR_A_T"], sd["HYDR_C_G"], sd["HYDR_G_T"])
            else:
                P = O.dna2_params(parse_temperature(T_STR), SALT)
            md = O.MD(P, c["pos"], O.axes_from_a1a3(c["a1"], c["a3"]), c["vel"], c["L"], t["btype"], t["n3"], t["n5"], c["box"], DT, 0.05)
            stepper = md.step
        barrier.wait()
        for _ in range(warm):
            stepper(md_steps)
        barrier.wait()
        for _ in range(steps):
            stepper(md_steps)
        barrier.wait()
        q.put("ok")
    except Exception as e:  # pragma: no cover
        q.put("error: " + repr(e))
        try:
            barrier.abort()
        except Exception:
            pass


def run_cpu(sysm, md_steps, warm, steps, procs, workload_name="c2", state=None):
    """Times `procs` independent single-threaded CPU simulations of the workload.  Returns (particle-steps/s, kind, seconds)."""
    import multiprocessing as mp
    from oracle import refharness
    from oracle import oracle as O
    kind = "reference" if refharness.available() else "port"
    if kind == "port":
        O.build()
    ctx = mp.get_context("spawn")
    d = tempfile.mkdtemp()
    top, conf = write_case(sysm, T_STR, d, state)
    keys = model_keys(workload_name, d)
    barrier = ctx.Barrier(procs + 1)
    q = ctx.Queue()
    ps = [ctx.Process(target=_ref_worker, args=(kind, top, conf, md_steps, warm, steps, barrier, q, keys)) for _ in range(procs)]
    for p in ps:
        p.start()
    barrier.wait()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    t1 = time.perf_counter()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    if any(r != "ok" for r in res):
        raise RuntimeError(str(res))
    N = len(sysm["pos"])
    return procs * N * md_steps * steps / (t1 - t0), kind, t1 - t0


def run_ref_cuda(sysm, workload_name, steps_a, steps_b, state, quick=False):
    """The reference's own CUDA backend on this GPU, same workload, via its stock CLI (oracle/ref_cuda_bench.py).
    It starts from `state`, a thermalised configuration produced by our engine: on the IDEAL lattice (exactly parallel
    base normals) the reference's GPU kernels normalise a zero cross product and fill the system with NaNs at step 0
    (profiles/micro/reflog.py), after which it never rebuilds its lists and its timing is void."""
    from oracle import ref_cuda_bench as R
    from oxdna_b200 import lattice
    if not R.available():
        return {"value": None, "unavailable": "oracle/_ref/oxDNA_cuda not built (make -f oracle/Makefile.refcuda)"}
    d = tempfile.mkdtemp()
    top, conf = write_case(sysm, T_STR, d, state)
    N = len(sysm["pos"])
    kw = dict(T=T_STR, salt=SALT, dt=DT)
    if workload_name == "c4":
        # the reference refuses external forces together with CUDA_sort_every > 0 (MD_CUDABackend.cu:110-112)
        variants = [(1, 0)] if quick else [(1, 0), (0, 0)]
        kw["ext_forces"] = lattice.mutual_traps(sysm)
    else:
        variants = [(1, 0)] if quick else ([(1, 1), (1, 0)] if workload_name == "c4_dna3" else [(1, 1), (0, 1), (1, 0), (0, 0)])
        kw["model_keys"] = model_keys(workload_name, d)
    res = R.time_reference_cuda(top, conf, N, steps_a, steps_b, variants, **kw)
    best = res["best"]
    out = {"value": best["value"] if best else None, "unit": "particle-steps/s", "best": best, "runs": res["runs"], "method": res["method"],
           "build": "unmodified /root/reference/src/CUDA, nvcc -arch=sm_100 -O3 -use_fast_math (oracle/Makefile.refcuda), backend_precision = mixed"}
    # second comparator: the same build with Timings.cpp compiled -DNOCUDA, i.e. WITHOUT a cudaDeviceSynchronize per timer
    # (src/Utilities/Timings.cpp:15-19,53-61) -- the reference's kernels and host logic with the timer synchronisation taken out
    if best and os.path.exists(R.BIN_NOSYNC):
        ns = R.time_reference_cuda(top, conf, N, steps_a, steps_b, [(best["use_edge"], best["CUDA_sort_every"])], binary=R.BIN_NOSYNC, **kw)
        out["no_timer_sync"] = {"value": ns["best"]["value"] if ns["best"] else None, "best": ns["best"], "method": ns["method"]}
        # MD_CUDABackend::sim_step reads the pinned flag _d_are_lists_old[0] right behind the launch of its first-step kernel
        # (src/CUDA/Backends/MD_CUDABackend.cu:570-581); the only thing that orders the read behind the kernel is the cudaDeviceSynchronize
        # of the timer that is paused in between (src/Utilities/Timings.cpp:53-61).  Without it the host reads a stale flag and the lists
        # are rebuilt late: such a run is faster and WRONG (pairs enter the cutoff unlisted).  The figure is kept as an upper bound only.
        try:
            ra, rb = ns["best"]["list_rebuild_every_md_steps"], best["list_rebuild_every_md_steps"]
            if ra > 2.0 * rb:
                out["no_timer_sync"]["valid"] = False
                out["no_timer_sync"]["note"] = (f"lists rebuilt every {ra:.0f} steps instead of every {rb:.1f}: without the timers' cudaDeviceSynchronize the reference "
                                                "reads its lists-are-old flag before the kernel has written it (MD_CUDABackend.cu:570-581) -- an upper bound from an "
                                                "incorrect run, not a comparator")
            else:
                out["no_timer_sync"]["valid"] = True
        except Exception:
            pass
    return out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload or ("c4" if args.gpus == 1 else "c2")
    sysm, desc = workload(wl)
    if args.gpus > 1:
        desc = "C5: 64 temperature replicas (290-350 K) of " + desc + " -- CPU arm: independent replicas of the same system, one per core"
    procs = args.ref_procs or min(host_cores(), 32)
    md = args.ref_md_steps or (1 if wl == "c4" else 4)
    val, kind, secs = run_cpu(sysm, md, args.warmup, args.steps, procs, wl)
    N = len(sysm["pos"])
    line = {"impl": "reference", "metric": "particle-steps/s", "value": val, "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "md_steps_per_step": md, "sample": f"{procs} independent single-threaded CPU processes x {md} MD steps per step",
                       "thermostat": "brownian", "dt": DT},
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": procs, "kind": kind,
                             "sample": f"{args.steps} x {md} MD steps of {N} nt on each of {procs} cores (aggregate)"},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


# ----------------------------------------------------------------------------------------------------------- our arm
def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    return float(peaks.get("hbm_gbs", 6650.0)), ("measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)")


# DRAM bytes of one force pass / one integrate launch from the committed `ncu --set full` capture of the same workload (cold cache: the
# compulsory traffic; profiles/summary_r02*.txt): dram__bytes_read.sum + dram__bytes_write.sum summed over the kernels of the pass
# (list: k_build_neigh half shell 247.3 MB + k_fill_edges 121.6 MB, ncu r02l; sort: the gather pass k_permute 407.7 MB, ncu r02f)
NCU_DRAM_BYTES = {"c2": dict(force=NCU_FORCE_PASS_DRAM_BYTES_C2, integrate=15.77e6),
                  "c4": dict(force=NCU_FORCE_PASS_DRAM_BYTES_C4, integrate=379.15e6, list=368.9e6, sort=407.7e6)}


def measure_single(args, wl, steps, warmup, equil, md, full, local_rank=0):
    """One system on one GPU: `steps` bench steps of `md` MD steps, CUDA events on the launching stream, device-side phase timeline
    (oxb_set_profile) over the SAME timed region.  full: add e2e, CPU baseline, reference-CUDA comparators."""
    import ctypes

    import torch
    from oxdna_b200 import capi, lattice
    from oxdna_b200.sim import Simulation, parse_temperature

    sysm, desc = workload(wl)
    N = len(sysm["pos"])
    v, L = lattice.maxwell_velocities(N, parse_temperature(T_STR), 5)
    conf = dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L)
    wargs = argparse.Namespace(**vars(args))
    wargs.workload = wl
    inp = base_input(wargs, T_STR)
    if wl == "c4":
        inp["external_forces_list"] = lattice.mutual_traps(sysm)
    sim = Simulation(inp, sysm, conf, device=local_rank)
    stream = torch.cuda.Stream()
    # the context launches on torch's stream so that torch.cuda.Event brackets exactly our kernels
    sim.ctx._ck(capi.lib().oxb_set_stream(sim.ctx._h, ctypes.c_void_p(stream.cuda_stream)))
    out = {}
    with torch.cuda.stream(stream):
        sim.run(equil)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        sim.ctx.set_profile(True)
        for _ in range(warmup):
            sim.run(md)
        sim.ctx.set_profile(True)  # zeroes the accumulators
        stats0, launches0 = sim.ctx.stats(), sim.ctx.launch_count()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        total_ms = 0.0
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sim.run(md)
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        clocks = sampler.stop()
        prof = sim.ctx.get_profile()
        sim.ctx.set_profile(False)
        launches = sim.ctx.launch_count() - launches0
        stats1 = sim.ctx.stats()
        n_md = md * steps
        n_reb = max(stats1["n_list_updates"] - stats0["n_list_updates"], 1)
        n_sort = max(stats1["n_sorts"] - stats0["n_sorts"], 1)
        value = N * n_md / (total_ms * 1e-3)
        step_ms = total_ms / n_md
        # ---- phase times from the device timeline of the timed region (they add up to it; launch gaps belong to the phase that was open)
        t_force, t_integ = prof["force"][0] / n_md, prof["integrate"][0] / n_md
        t_build, t_sort, t_wait = (prof["build"][0] + prof["edges"][0]) / n_reb, (prof["sort"][0] + prof["permute"][0]) / n_sort, prof["wait"][0] / n_reb
        t_parts = {"sort_keys_radix_invert": prof["sort"][0] / n_sort, "sort_gather_pass": prof["permute"][0] / n_sort,
                   "build_neighbour_scan": prof["build"][0] / n_reb, "build_edge_scan_fill": prof["edges"][0] / n_reb}
        prof_total = sum(v[0] for v in prof.values())
        ps = pair_statistics(sim, sysm)
        hbm_peak, peak_src = load_peaks()
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        flops = N * (FLOP_FAR * ps["listed"] + FLOP_DH * ps["dh"] + FLOP_CONTACT * ps["contact"] + FLOP_BONDED * 1.0)
        particle_centric = not args.use_edge
        if particle_centric:
            flops = N * (2 * (FLOP_FAR * ps["listed"] + FLOP_DH * ps["dh"] + FLOP_CONTACT * ps["contact"]) + 2 * FLOP_BONDED)
        tf = flops / (t_force * 1e-3) / 1e12
        # algorithmic bytes: force pass 40 B state read + 32 B F,T write + 8 B per listed unique pair; integrate 336 B (+ 16 B
        # Debye-Hueckel site force in edge mode); list rebuild 140 B per particle; Hilbert sort 470 B per particle
        force_bytes = N * (40.0 + 32.0 + 8.0 * ps["listed"])
        integ_bytes = N * (336.0 + (16.0 if args.use_edge else 0.0))
        gbs = lambda b, ms: b / (ms * 1e-3) / 1e9 if ms > 0 else None
        ncu = NCU_DRAM_BYTES.get(wl, {}) if args.use_edge else {}
        out.update({
            "value": value, "ms_per_step": total_ms / steps, "gpu_launches": int(launches), "clocks": clocks, "N": N, "desc": desc, "pairs_per_particle": ps,
            "step_ms": step_ms, "list_rebuild_every_md_steps": n_md / n_reb,
            "roofline": {"kernel": "force pass (Debye-Hueckel + near edges + HB/cross stacking + coaxial + bonded)" if not particle_centric else "forces (particle-centric)",
                         "bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak, "traffic": ncu.get("force"),
                         "model": "SURVEY 8(d) algorithmic FLOP per pair class x measured pair counts", "sm_mhz": sm_mhz, "ms": t_force,
                         "share_of_step": t_force / step_ms,
                         "hbm": {"achieved": gbs(force_bytes, t_force), "peak": hbm_peak, "unit": "GB/s", "frac": gbs(force_bytes, t_force) / hbm_peak,
                                 "algorithmic_bytes": force_bytes, "peak_source": peak_src},
                         "timing": "device-side %globaltimer phase timeline inside the timed region (oxb_set_profile)"},
            "roofline_integrate": {"kernel": "k_integrate: second half-kick + thermostat + first half-kick/drift/rotate + staleness", "bound": "hbm",
                                   "achieved": gbs(integ_bytes, t_integ), "peak": hbm_peak, "unit": "GB/s", "frac": gbs(integ_bytes, t_integ) / hbm_peak,
                                   "traffic": ncu.get("integrate"), "ms": t_integ, "share_of_step": t_integ / step_ms, "peak_source": peak_src},
            "roofline_list": {"kernel": "list rebuild (binning + 27-cell scan + DH matrix + edge list)", "bound": "hbm", "achieved": gbs(N * 140.0, t_build),
                              "peak": hbm_peak, "unit": "GB/s", "frac": gbs(N * 140.0, t_build) / hbm_peak, "traffic": ncu.get("list"), "ms": t_build,
                              "per_md_step_ms": t_build * n_reb / n_md, "note": "issue-bound on divergence, not on bytes (DESIGN 3)"},
            "roofline_sort": {"kernel": "Hilbert re-sort (keys + radix sort + one gather pass)", "bound": "hbm", "achieved": gbs(N * 470.0, t_sort),
                              "peak": hbm_peak, "unit": "GB/s", "frac": gbs(N * 470.0, t_sort) / hbm_peak if t_sort > 0 else None, "traffic": ncu.get("sort"), "ms": t_sort,
                              "per_md_step_ms": t_sort * n_sort / n_md},
            "kernels_ms": {"force_pass": t_force, "integrate": t_integ, "list_build_per_rebuild": t_build, "sort_per_sort": t_sort, "rebuild_parts": t_parts,
                           "halt_and_host_wait_per_rebuild": t_wait, "batch_launch_gap_per_batch": prof["gap"][0] / max(prof["gap"][1], 1), "batches": prof["gap"][1],
                           "batch_launch_gap_per_md_step": prof["gap"][0] / n_md, "other_per_md_step": prof["other"][0] / n_md, "md_step_mean": step_ms,
                           "sum_of_phases_per_md_step": prof_total / n_md, "rebuilds": n_reb, "sorts": n_sort,
                           "source": "device-side phase timeline of the timed region; phases add up to md_step_mean"},
        })
        if not full:
            if args.extras_ref_cuda:
                try:
                    out["reference_cuda"] = run_ref_cuda(sysm, wl, 5000, 15000, sim.ctx.get_state(), quick=True)
                except Exception as e:  # pragma: no cover
                    out["reference_cuda"] = {"value": None, "unavailable": repr(e)[-300:]}
            sim.close()
            return out

        # ---- end to end through the public API with host buffers (OxpyManager.run semantics: H2D, steps, D2H)
        st = sim.ctx.get_state()
        pin = {k: torch.from_numpy(st[k]).pin_memory().numpy() for k in ("pos", "a1", "a3", "vel", "L")}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(steps, 3))
        for _ in range(e2e_steps):
            sim.ctx.set_state(pin["pos"], pin["a1"], pin["a3"], pin["vel"], pin["L"])
            sim.run(md)
            sim.ctx.get_state(out=pin)  # D2H straight into the pinned host buffers
            U, K = sim.ctx.energy()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        out["e2e"] = {"value": N * md / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": N * 120, "d2h_bytes_per_step": N * 120 + 16,
                      "api": "set_state (H2D from pinned host buffers) + run(md_steps) + get_state + energy (D2H), host wall clock", "U_per_particle": U / N,
                      "K_per_particle": K / N}
        state = sim.ctx.get_state()
        sim.close()
        del flush
        torch.cuda.empty_cache()

    # ---- CPU baseline on a bounded sample of the same workload
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cmd = args.cpu_md_steps or (6 if wl == "c4" else 40)
            cval, kind, secs = run_cpu(sysm, cmd, 0, 1, 1, wl, state)
            cpu = {"value": cval, "unit": "particle-steps/s", "cores": 1, "kind": kind,
                   "sample": f"{cmd} MD steps of the same {N}-nt system on one host core ({secs:.1f} s), host has {host_cores()} cores"}
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": "particle-steps/s", "cores": 1, "kind": "port", "sample": "failed: " + repr(e)}
    out["cpu_baseline"] = cpu
    ref_cuda = None
    if not args.no_ref_cuda:
        try:
            a, b = args.ref_cuda_steps or ([3000, 9000] if wl == "c4" else ([2000, 5000] if wl == "c4_dna3" else [10000, 20000]))
            ref_cuda = run_ref_cuda(sysm, wl, a, b, state)
        except Exception as e:  # pragma: no cover
            ref_cuda = {"value": None, "unavailable": repr(e)[-300:]}
    out["reference_cuda"] = ref_cuda
    return out


def measure_ensemble(args, R_total, steps, warmup, equil, md, world, rank, local_rank, dist=None, e2e=True):
    """Config C5: R_total temperature replicas of C2 (geometric 290-350 K), R_total / world per GPU held as replica BATCHES (one context,
    one launch per kernel for all local replicas), one exchange attempt (NCCL all_gather of 2 doubles per replica) per bench step."""
    import torch
    from oxdna_b200 import lattice
    from oxdna_b200.remd import ReplicaExchange, TorchComm, geometric_ladder
    from oxdna_b200.sim import make_batches, parse_temperature

    sysm, desc = workload("c2")
    N = len(sysm["pos"])
    nl = R_total // world
    ladder_K = geometric_ladder(290.0, 350.0, R_total)
    ladder = ladder_K * 0.1 / 300.0
    confs = []
    for k in range(nl):
        g = rank * nl + k
        v, L = lattice.maxwell_velocities(N, float(ladder[g]), 5 + g)
        confs.append(dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L))
    wargs = argparse.Namespace(**vars(args))
    wargs.workload = "c2"
    inp = base_input(wargs, T_STR)
    inp.pop("T")
    batches = make_batches(inp, sysm, confs, ladder[rank * nl:(rank + 1) * nl], device=local_rank, ladder_max=float(ladder[-1]))
    comm = TorchComm(torch.device("cuda", local_rank)) if world > 1 else None
    remd = ReplicaExchange(batches, ladder, comm, seed=42)
    remd.advance(equil)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        remd.advance(md)
        remd.exchange()
    launches0 = sum(b.ctx.launch_count() for b in batches)
    stats0 = [b.ctx.stats() for b in batches]
    for k in remd.timers:
        remd.timers[k] = 0.0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = 0.0
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        remd.advance(md)
        remd.exchange()
        torch.cuda.synchronize()  # the batches run on their own streams: bracket with device-wide synchronisation
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    my_ms = total_ms
    timers = dict(remd.timers)
    launches = sum(b.ctx.launch_count() for b in batches) - launches0
    stats1 = [b.ctx.stats() for b in batches]
    rebuilds = sum(s1["n_list_updates"] - s0["n_list_updates"] for s0, s1 in zip(stats0, stats1))
    # ---- end to end at N GPUs: every bench step uploads the local replicas' states from pinned host memory, advances, exchanges, reads back
    e2e_ms = None
    if e2e:
        pins = []
        for b in batches:
            st = b.ctx.get_state()
            pins.append({k: torch.from_numpy(st[k]).pin_memory().numpy() for k in ("pos", "a1", "a3", "vel", "L")})
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(steps, 3))
        for _ in range(n_e2e):
            for b, pin in zip(batches, pins):
                b.ctx.set_state(pin["pos"], pin["a1"], pin["a3"], pin["vel"], pin["L"])
            remd.advance(md)
            remd.exchange()
            for b, pin in zip(batches, pins):
                b.ctx.get_state(out=pin)
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    per_rank = [dict(rank=rank, device_ms_per_step=my_ms / steps, md_ms=1e3 * timers["md"] / steps, exchange_ms=1e3 * (timers["energy"] + timers["update"]) / steps,
                     host_wait_ms=1e3 * timers["comm"] / steps, list_rebuilds_per_replica_batch=rebuilds / max(len(batches), 1) / steps)]
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_max = float(t[0].item()), float(t[1].item())
        e2e_ms = e2e_max if e2e else None
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank[0])
        per_rank = gathered
        la = torch.tensor([launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(la)
        launches = int(la.item())
    value = R_total * N * md * steps / (total_ms * 1e-3)
    res = {"value": value, "ms_per_step": total_ms / steps, "gpu_launches": int(launches), "clocks": clocks, "per_rank": per_rank,
           "config": {"workload": "C5: replica-exchange MD, %d temperature replicas (geometric 290-350 K) of " % R_total + desc, "md_steps_per_step": md,
                      "replicas": R_total, "replicas_per_gpu": nl, "batches_per_gpu": len(batches),
                      "parallelism": f"{nl} replicas per GPU as {len(batches)} replica batch(es): one context, one launch per kernel for all replicas of a batch "
                                     "(replica = grid offset, per-replica constant table); replica exchange (temperature swap) every bench step, NCCL all_gather of "
                                     "2 doubles per replica; no data-path collective",
                      "ladder_K": [float(ladder_K[0]), float(ladder_K[-1])], "use_edge": int(args.use_edge), "CUDA_sort_every": args.sort_every,
                      "thermostat": "brownian (newtonian_steps 103)", "dt": DT, "verlet_skin": 0.05, "equilibration_md_steps": equil,
                      "l2": "256 MiB buffer written between timed iterations (L2 flush)", "exchange_acceptance": float(np.mean(remd.rates()))}}
    if e2e_ms:
        res["e2e"] = {"value": R_total * N * md / (e2e_ms * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": R_total * N * 120,
                      "d2h_bytes_per_step": R_total * N * 120 + 16 * R_total, "n_gpus": world,
                      "api": "per bench step on every rank: set_state of the local replica batches (H2D from pinned host buffers) + run(md_steps) + exchange "
                             "(energies D2H, NCCL all_gather) + get_state (D2H); host wall clock, max over ranks"}
    for b in batches:
        b.close()
    del flush
    torch.cuda.empty_cache()
    return res


def ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: oxdna_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    base = {"metric": "particle-steps/s", "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "vs_baseline": None, "dtype": "f32 forces / f64 integration (mixed)", "data": "synthetic"}
    if world > 1 or args.workload == "c5":
        R = args.replicas or 64
        if R % world:
            raise ValueError(f"{R} replicas do not divide over {world} GPUs")
        r = measure_ensemble(args, R, args.steps, args.warmup, args.equil, args.md_steps, world, rank, local_rank, dist)
        if rank == 0:
            line = dict(base, value=r["value"], ms_per_step=r["ms_per_step"], scaling="strong", config=r["config"], gpu_launches=r["gpu_launches"], clocks=r["clocks"],
                        e2e=r.get("e2e"), per_rank=r["per_rank"], roofline=None, cpu_baseline=None,
                        note="strong scaling: the 64 replicas of C5 are divided over the GPUs; the single-GPU C5 figure (64 replicas on one GPU) is extras.c5_one_gpu "
                             "of the --gpus 1 line, whose headline is C4")
            print(json.dumps(line), file=_REAL_STDOUT, flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    wl = args.workload or "c4"
    m = measure_single(args, wl, args.steps, args.warmup, args.equil, args.md_steps, True, local_rank)
    ps, step_ms = m.pop("pairs_per_particle"), m.pop("step_ms")
    line = dict(base, value=m.pop("value"), ms_per_step=m.pop("ms_per_step"), scaling="weak",
                config={"workload": m.pop("desc"), "md_steps_per_step": args.md_steps, "replicas": 1, "parallelism": "single system (a single system does not shard: replicas only)",
                        "use_edge": int(args.use_edge), "CUDA_sort_every": args.sort_every, "thermostat": "brownian (newtonian_steps 103)", "dt": DT, "verlet_skin": 0.05,
                        "equilibration_md_steps": args.equil, "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                        "ns_per_day": 86400.0 / (step_ms * 1e-3) * DT * 3.03e-3, "md_steps_per_s": 1e3 / step_ms,
                        "list_rebuild_every_md_steps": m.pop("list_rebuild_every_md_steps"), "pairs_per_particle": ps})
    m.pop("N")
    line.update(m)
    # ---- the other BASELINE configs, bounded: C2 and C3 (single systems) and C5 on ONE GPU (64 replicas, two batches)
    extras = {}
    if not args.no_extras and wl == "c4":
        for name in ("c2", "c3", "c2_dna3"):
            try:
                x = measure_single(args, name, 5, 3, 10000, 1000, False, local_rank)
                extras[name] = {"workload": x["desc"], "value": x["value"], "unit": "particle-steps/s", "md_step_ms": x["step_ms"], "kernels_ms": x["kernels_ms"],
                                "roofline_fp32_frac": x["roofline"]["frac"], "list_rebuild_every_md_steps": x["list_rebuild_every_md_steps"],
                                "reference_cuda": x.get("reference_cuda")}
            except Exception as e:  # pragma: no cover
                extras[name] = {"error": repr(e)[-300:]}
        try:
            x = measure_ensemble(args, args.replicas or 64, 3, 2, 3000, 1000, 1, 0, local_rank, None, e2e=False)
            extras["c5_one_gpu"] = {"workload": x["config"]["workload"], "value": x["value"], "unit": "particle-steps/s", "ms_per_step": x["ms_per_step"],
                                    "batches": x["config"]["batches_per_gpu"], "per_rank": x["per_rank"], "exchange_acceptance": x["config"]["exchange_acceptance"],
                                    "note": "the strong-scaling base of the --gpus N runs (64 replicas on one GPU)"}
        except Exception as e:  # pragma: no cover
            extras["c5_one_gpu"] = {"error": repr(e)[-300:]}
    line["extras"] = extras
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--workload", default=None, choices=["c2", "c3", "c4", "c5", "small", "c2_dna3", "c4_dna3"], help="default: c4 on one GPU, c5 (replica ensemble) on several")
    ap.add_argument("--md-steps", type=int, default=1000, help="MD steps per bench step")
    ap.add_argument("--equil", type=int, default=10000, help="untimed equilibration MD steps")
    ap.add_argument("--use-edge", type=int, default=1)
    ap.add_argument("--sort-every", type=int, default=1)
    ap.add_argument("--replicas", type=int, default=0, help="C5: total number of temperature replicas (default 64)")
    ap.add_argument("--ref-md-steps", type=int, default=0, help="MD steps per bench step of the CPU reference arm (default: 1 at C4, 4 otherwise)")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--cpu-md-steps", type=int, default=0, help="MD steps of the one-core CPU baseline sample (default: 6 at C4, 40 otherwise)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-CUDA-backend comparator legs")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the bounded C2 / C3 / C5-on-one-GPU legs")
    ap.add_argument("--extras-ref-cuda", type=int, default=1, help="reference-CUDA comparator (one variant) for the C2 / C3 extras")
    ap.add_argument("--ref-cuda-steps", type=int, nargs=2, default=None, help="steps=A and steps=B runs of the reference CLI")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that print banners there (NCCL's version line at communicator creation, torchrun
    # notices) are sent to stderr for the duration of the run; json lines go to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args)
    elif args.impl == "reference-cuda":
        if int(os.environ.get("RANK", "0")) == 0:
            from oxdna_b200 import lattice
            from oxdna_b200.sim import Simulation, parse_temperature
            wl = args.workload or "c4"
            sysm, desc = workload(wl)
            a, b = args.ref_cuda_steps or [10000, 20000]
            v, L = lattice.maxwell_velocities(len(sysm["pos"]), parse_temperature(T_STR), 5)
            wargs = argparse.Namespace(**vars(args))
            wargs.workload = wl
            sim = Simulation(base_input(wargs), sysm, dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L))
            sim.run(args.equil)
            state = sim.ctx.get_state()
            sim.close()
            r = run_ref_cuda(sysm, wl, a, b, state)
            print(json.dumps({"impl": "reference-cuda", "metric": "particle-steps/s", "value": r.get("value"), "unit": "particle-steps/s", "n_gpus": 1,
                              "higher_is_better": True, "config": {"workload": desc}, "reference_cuda": r}), file=_REAL_STDOUT, flush=True)
    else:
        ours(args)


if __name__ == "__main__":
    main()
