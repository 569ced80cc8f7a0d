#!/usr/bin/env python
"""Benchmark of the oxDNA GPU MD step (BASELINE.json metric: particle-steps/s).

  python bench.py --gpus 1 --steps K --warmup W              our CUDA path, N = 1: config C2 (81,920 nt duplex lattice)
  torchrun ... bench.py --gpus N ...                          N > 1: one C2 replica per GPU, replica exchange over NCCL
  python bench.py --impl reference ...                        the reference's own CPU implementation on the host cores

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
# ncu dram__bytes_read.sum + dram__bytes_write.sum of one force pass (near + HB/CRST + coaxial + bonded + DH kernels)
# (profiles/summary_r01n.txt, one `ncu --set full` capture per workload, cold cache: compulsory traffic of one pass;
#  kernels: Debye-Hueckel, bonded, near edges, HB / cross stacking, coaxial stacking, excluded volume in double)
NCU_FORCE_PASS_DRAM_BYTES_C2 = int((6.92 + 7.60 + 4.77 + 6.64 + 0.03 + 0.38) * 1e6)
NCU_FORCE_PASS_DRAM_BYTES_C4 = int((88.77 + 92.25 + 59.75 + 83.87 + 0.03 + 4.03) * 1e6)


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
    elif name == "small":
        sysm = lattice.duplex_lattice(64, bp=20, spacing=10.0, seed=12345)
        desc = "small: 64 x 20-bp duplexes (2,560 nt)"
    else:
        raise ValueError(name)
    return sysm, desc


def model_keys(name, tmpdir=None):
    """interaction keys of the workload; with tmpdir the sequence-dependent table is written to a file (reference binaries)"""
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


def run_ref_cuda(sysm, workload_name, steps_a, steps_b, state):
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
    if workload_name == "c4":
        # the reference refuses external forces together with CUDA_sort_every > 0 (MD_CUDABackend.cu:110-112)
        res = R.time_reference_cuda(top, conf, N, steps_a, steps_b, [(1, 0), (0, 0)], T=T_STR, salt=SALT, dt=DT, ext_forces=lattice.mutual_traps(sysm))
    else:
        res = R.time_reference_cuda(top, conf, N, steps_a, steps_b, [(1, 1), (0, 1), (1, 0), (0, 0)], T=T_STR, salt=SALT, dt=DT,
                                    model_keys=model_keys(workload_name, d))
    best = res["best"]
    return {"value": best["value"] if best else None, "unit": "particle-steps/s", "best": best, "runs": res["runs"], "method": res["method"],
            "build": "unmodified /root/reference/src/CUDA, nvcc -arch=sm_100 -O3 -use_fast_math (oracle/Makefile.refcuda), backend_precision = mixed"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sysm, desc = workload(args.workload)
    procs = args.ref_procs or min(host_cores(), 32)
    md = args.ref_md_steps
    val, kind, secs = run_cpu(sysm, md, args.warmup, args.steps, procs, args.workload)
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
def ours(args):
    import torch
    import torch.distributed as dist
    from oxdna_b200 import capi, lattice
    from oxdna_b200.remd import ReplicaExchange, TorchComm, geometric_ladder
    from oxdna_b200.sim import Simulation, parse_temperature

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: oxdna_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sysm, desc = workload(args.workload)
    N = len(sysm["pos"])
    md = args.md_steps
    n_rep = world  # one replica per GPU: weak scaling
    # default scaling run: one C2 replica per GPU on a NARROW ladder around the 300 K of the single-GPU workload (geometric 299-301 K:
    # ~0.3 K spacing at 8 replicas, where 81,920-nt replicas still exchange), so that every rank carries the same work as the N = 1
    # run; the 290-350 K ladder of config C5 is what --replicas-per-gpu runs (hotter replicas rebuild their lists more often)
    ladder_K = geometric_ladder(299.0, 301.0, n_rep) if n_rep > 1 else np.array([300.0])
    T_sim = ladder_K * 0.1 / 300.0
    myT = f"{ladder_K[rank]:.6f}K"
    v, L = lattice.maxwell_velocities(N, parse_temperature(myT), 5 + rank)
    conf = dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L)
    inp = base_input(args, myT)
    if args.workload == "c4":
        inp["external_forces_list"] = lattice.mutual_traps(sysm)
    sim = Simulation(inp, sysm, conf, device=local_rank)
    stream = torch.cuda.Stream()
    # the context launches on torch's stream so that torch.cuda.Event brackets exactly our kernels
    sim.ctx._ck(capi.lib().oxb_set_stream(sim.ctx._h, __import__("ctypes").c_void_p(stream.cuda_stream)))
    remd = ReplicaExchange([sim], T_sim, TorchComm(torch.device("cuda", local_rank)) if world > 1 else None, seed=42) if world > 1 else None

    with torch.cuda.stream(stream):
        sim.run(args.equil)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

        def one_step():
            sim.run(md)
            if remd is not None:
                remd.exchange()

        for _ in range(args.warmup):
            one_step()
        stats0 = sim.ctx.stats()
        launches0 = sim.ctx.launch_count()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        total_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (state of C2 is smaller than the 126 MB L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            one_step()
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            dist.barrier()
            t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        launches = sim.ctx.launch_count() - launches0
        stats1 = sim.ctx.stats()
        value = n_rep * N * md * args.steps / (total_ms * 1e-3)

        if rank != 0:
            if world > 1:
                dist.barrier()
                dist.destroy_process_group()
            return

        # ---- per-kernel roofline figures, measured live with CUDA events on the launching stream
        t_force = sim.ctx.time_kernel(0, 20)
        t_integ = sim.ctx.time_kernel(1, 20)
        t_list = sim.ctx.time_kernel(2, 5)
        t_sort = sim.ctx.time_kernel(3, 5)
        ps = pair_statistics(sim, sysm)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # algorithmic bytes per particle of the force pass: 40 B state read + 32 B F,T write + 8 B per listed unique pair
        force_bytes = N * (40.0 + 32.0 + 8.0 * ps["listed"])
        force_gbs = force_bytes / (t_force * 1e-3) / 1e9
        flops = N * (FLOP_FAR * ps["listed"] + FLOP_DH * ps["dh"] + FLOP_CONTACT * ps["contact"] + FLOP_BONDED * 1.0)
        if not args.use_edge:
            flops = N * (2 * (FLOP_FAR * ps["listed"] + FLOP_DH * ps["dh"] + FLOP_CONTACT * ps["contact"]) + 2 * FLOP_BONDED)
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        # first_step_mixed: 176 B read + 160 B written per particle, second half-kick fused; + 16 B Fb read in edge mode
        integ_bytes = N * (336.0 + (16.0 if args.use_edge else 0.0))
        integ_gbs = integ_bytes / (t_integ * 1e-3) / 1e9
        # DRAM bytes of one force pass from the committed ncu --set full capture of the same workload (sum over its kernels)
        traffic = {"c2": NCU_FORCE_PASS_DRAM_BYTES_C2, "c4": NCU_FORCE_PASS_DRAM_BYTES_C4}.get(args.workload) if args.use_edge else None
        step_ms = total_ms / (args.steps * md)
        rebuild_every = md * args.steps / max(stats1["n_list_updates"] - stats0["n_list_updates"], 1)

        # ---- end to end through the public API with host buffers (OxpyManager.run semantics: H2D, steps, D2H)
        st = sim.ctx.get_state()
        pin = {k: torch.from_numpy(st[k]).pin_memory().numpy() for k in ("pos", "a1", "a3", "vel", "L")}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            sim.ctx.set_state(pin["pos"], pin["a1"], pin["a3"], pin["vel"], pin["L"])
            sim.run(md)
            sim.ctx.get_state(out=pin)  # D2H straight into the pinned host buffers
            U, K = sim.ctx.energy()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        e2e = {"value": N * md / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": N * 172, "d2h_bytes_per_step": N * 144 + 16,
               "api": "set_state (H2D) + run(md_steps) + get_state + energy (D2H), wall clock, n_gpus=1 leg", "U_per_particle": U / N, "K_per_particle": K / N}

        # ---- CPU baseline on a bounded sample of the same workload (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cval, kind, secs = run_cpu(sysm, args.cpu_md_steps, 0, 1, 1, args.workload, sim.ctx.get_state())
                cpu = {"value": cval, "unit": "particle-steps/s", "cores": 1, "kind": kind,
                       "sample": f"{args.cpu_md_steps} MD steps of the same {N}-nt system on one host core ({secs:.1f} s), host has {host_cores()} cores"}
            except Exception as e:  # pragma: no cover
                cpu = {"value": None, "unit": "particle-steps/s", "cores": 1, "kind": "port", "sample": "failed: " + repr(e)}

        ref_cuda = None
        if world == 1 and not args.no_ref_cuda:
            try:
                a, b = args.ref_cuda_steps
                ref_cuda = run_ref_cuda(sysm, args.workload, a, b, sim.ctx.get_state())
            except Exception as e:  # pragma: no cover
                ref_cuda = {"value": None, "unavailable": repr(e)[-300:]}

        line = {"metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 forces / f64 integration (mixed)",
                "data": "synthetic",
                "config": {"workload": desc, "md_steps_per_step": md, "replicas": n_rep, "parallelism": (f"1 replica per GPU, temperatures geometric {float(ladder_K[0]):.1f}-{float(ladder_K[-1]):.1f} K, replica-exchange attempt (NCCL all_gather of "
                                           "2 doubles per replica) every bench step") if world > 1 else "single system",
                           "use_edge": int(args.use_edge), "CUDA_sort_every": args.sort_every, "thermostat": "brownian (newtonian_steps 103)", "dt": DT, "verlet_skin": 0.05,
                           "equilibration_md_steps": args.equil, "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                           "ns_per_day": 86400.0 / (step_ms * 1e-3) * DT * 3.03e-3, "md_steps_per_s": 1e3 / step_ms, "list_rebuild_every_md_steps": rebuild_every,
                           "pairs_per_particle": ps},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"kernel": "forces (edge non-bonded + bonded)" if args.use_edge else "forces (particle-centric)", "bound": "hbm", "achieved": force_gbs, "peak": hbm_peak,
                             "unit": "GB/s", "frac": force_gbs / hbm_peak, "traffic": traffic,
                             "traffic_source": "profiles/summary_r01n.txt: dram__bytes_read.sum + dram__bytes_write.sum summed over the kernels of one force pass", "peak_source": peak_src, "ms": t_force,
                             "share_of_step": t_force / step_ms,
                             "note": "the force kernel is FP32/SFU-bound, not HBM-bound (see roofline_fp32); HBM figure given as the contract asks"},
                "roofline_fp32": {"achieved_tflops": flops / (t_force * 1e-3) / 1e12, "peak_tflops": fp32_peak, "frac": flops / (t_force * 1e-3) / 1e12 / fp32_peak,
                                  "model": "SURVEY 8(d) algorithmic FLOP per pair class", "sm_mhz": sm_mhz},
                "roofline_integrate": {"kernel": "fused second half-kick + thermostat + first half-kick/drift/rotate", "bound": "hbm", "achieved": integ_gbs, "peak": hbm_peak,
                                       "unit": "GB/s", "frac": integ_gbs / hbm_peak, "ms": t_integ, "share_of_step": t_integ / step_ms},
                "kernels_ms": {"forces": t_force, "integrate": t_integ, "rebuild_incl_sort": t_list, "sort_only": t_sort, "md_step_mean": step_ms,
                               "rebuild_amortised": t_list / rebuild_every},
                "cpu_baseline": cpu, "reference_cuda": ref_cuda}
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()


def ours_ensemble(args):
    """Config C5: `--replicas-per-gpu` temperature replicas of the workload on every GPU (64 in total at 8 x 8), advanced
    concurrently (one host thread and one set of CUDA streams per replica), one exchange attempt per bench step."""
    import torch
    import torch.distributed as dist
    from oxdna_b200 import lattice
    from oxdna_b200.remd import ReplicaExchange, TorchComm, geometric_ladder
    from oxdna_b200.sim import Simulation, parse_temperature

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: oxdna_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sysm, desc = workload(args.workload)
    N, md, nl = len(sysm["pos"]), args.md_steps, args.replicas_per_gpu
    R = world * nl
    ladder_K = geometric_ladder(290.0, 350.0, R)
    sims = []
    for k in range(nl):
        g = rank * nl + k
        T = f"{ladder_K[g]:.6f}K"
        v, L = lattice.maxwell_velocities(N, parse_temperature(T), 5 + g)
        sims.append(Simulation(base_input(args, T), sysm, dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L), device=local_rank))
    remd = ReplicaExchange(sims, ladder_K * 0.1 / 300.0, TorchComm(torch.device("cuda", local_rank)) if world > 1 else None, seed=42,
                           concurrent=not args.sequential_replicas)
    remd.advance(args.equil)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(args.warmup):
        remd.advance(md)
        remd.exchange()
    launches0 = sum(s.ctx.launch_count() for s in sims)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        remd.advance(md)
        remd.exchange()
        torch.cuda.synchronize()  # the replicas run on their own streams: bracket with device-wide synchronisation
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.barrier()
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    launches = sum(s.ctx.launch_count() for s in sims) - launches0
    if rank == 0:
        value = R * N * md * args.steps / (total_ms * 1e-3)
        line = {"metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 forces / f64 integration (mixed)", "data": "synthetic",
                "config": {"workload": "C5-style ensemble: " + desc, "md_steps_per_step": md, "replicas": R, "replicas_per_gpu": nl,
                           "parallelism": f"{nl} replicas per GPU advanced {'sequentially' if args.sequential_replicas else 'concurrently (one host thread + CUDA streams each)'}, "
                                          "replica exchange (temperature swap) every bench step, NCCL all_gather of 2 doubles per replica",
                           "ladder_K": [float(ladder_K[0]), float(ladder_K[-1])], "use_edge": int(args.use_edge), "CUDA_sort_every": args.sort_every,
                           "l2": "256 MiB buffer written between timed iterations", "exchange_acceptance": float(np.mean(remd.rates()))},
                "gpu_launches": int(launches), "clocks": clocks, "e2e": None, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "small"])
    ap.add_argument("--md-steps", type=int, default=1000, help="MD steps per bench step")
    ap.add_argument("--equil", type=int, default=10000, help="untimed equilibration MD steps")
    ap.add_argument("--use-edge", type=int, default=1)
    ap.add_argument("--sort-every", type=int, default=1)
    ap.add_argument("--ref-md-steps", type=int, default=4, help="MD steps per bench step of the CPU reference arm")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--cpu-md-steps", type=int, default=40)
    ap.add_argument("--replicas-per-gpu", type=int, default=1, help="> 1: C5-style replica ensemble (64 replicas = 8 per GPU on 8 GPUs)")
    ap.add_argument("--sequential-replicas", action="store_true", help="ensemble mode: advance the local replicas one after the other (comparison)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-CUDA-backend comparator leg")
    ap.add_argument("--ref-cuda-steps", type=int, nargs=2, default=[10000, 20000], help="steps=A and steps=B runs of the reference CLI")
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
            sysm, desc = workload(args.workload)
            a, b = args.ref_cuda_steps
            v, L = lattice.maxwell_velocities(len(sysm["pos"]), parse_temperature(T_STR), 5)
            sim = Simulation(base_input(args), sysm, dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L))
            sim.run(args.equil)
            state = sim.ctx.get_state()
            sim.close()
            r = run_ref_cuda(sysm, args.workload, a, b, state)
            print(json.dumps({"impl": "reference-cuda", "metric": "particle-steps/s", "value": r.get("value"), "unit": "particle-steps/s", "n_gpus": 1,
                              "higher_is_better": True, "config": {"workload": desc}, "reference_cuda": r}), file=_REAL_STDOUT, flush=True)
    elif args.replicas_per_gpu > 1:
        ours_ensemble(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
