"""BENCH INFRASTRUCTURE ONLY.  Times the UNMODIFIED reference CUDA backend (oracle/_ref/oxDNA_cuda, built by
oracle/Makefile.refcuda from /root/reference/src/CUDA) on the same synthetic workload bench.py uses, through the
reference's own CLI and input file -- the comparator BASELINE.json names ("the reference's own CUDA backend on one B200").

Method: the stock binary is run twice from the same files with `steps = A` and `steps = B` (A < B); the difference of
the reference's own "Total Running Time" (its SimBackend timer, simulation loop only) covers MD steps A..B, i.e. the
same post-equilibration region bench.py times for our path.  Input follows SURVEY.md appendix C (timers left on, as users run it).
The start configuration must be thermalised (velocities included, `refresh_vel = 0`): see bench.run_ref_cuda.
Never imported by the product package.
"""
import os
import re
import subprocess
import tempfile
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(_HERE, "_ref", "oxDNA_cuda")
# same objects with Timings.cpp compiled -DNOCUDA: no cudaDeviceSynchronize per timer (src/Utilities/Timings.cpp:15-19,53-61)
BIN_NOSYNC = os.path.join(_HERE, "_ref", "oxDNA_cuda_nosync")


def available():
    return os.path.exists(BIN)


TEMPLATE = """backend = CUDA
backend_precision = mixed
sim_type = MD
{model}
salt_concentration = {salt}
T = {T}
dt = {dt}
steps = {steps}
thermostat = brownian
newtonian_steps = 103
diff_coeff = 2.5
verlet_skin = 0.05
CUDA_list = verlet
CUDA_sort_every = {sort_every}
use_edge = {use_edge}
edge_n_forces = 1
max_density_multiplier = 3
CUDA_avoid_cpu_calculations = 1
seed = 42
refresh_vel = 0
topology = {top}
conf_file = {conf}
trajectory_file = {d}/trajectory.dat
lastconf_file = {d}/last_conf.dat
energy_file = {d}/energy.dat
log_file = {d}/log.dat
print_energy_every = 100000000
print_conf_interval = 100000000
restart_step_counter = 1
time_scale = linear
no_stdout_energy = 1
external_forces = {ext}
{extfile}
"""


def write_forces_file(path, forces):
    with open(path, "w") as f:
        for e in forces:
            f.write("{\n")
            for k, v in e.items():
                f.write(f"{k} = {v}\n")
            f.write("}\n")


def _run(d, top, conf, steps, use_edge, sort_every, T, salt, dt, ext_path, model_keys=None, binary=BIN):
    inp = os.path.join(d, f"input_{steps}")
    with open(inp, "w") as f:
        model = "\n".join(f"{k} = {v}" for k, v in (model_keys or {"interaction_type": "DNA2"}).items())
        f.write(TEMPLATE.format(model=model, salt=salt, T=T, dt=dt, steps=steps, sort_every=sort_every, use_edge=use_edge, top=top, conf=conf, d=d,
                                ext=1 if ext_path else 0, extfile=f"external_forces_file = {ext_path}" if ext_path else ""))
    t0 = time.perf_counter()
    p = subprocess.run([binary, inp], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    t1 = time.perf_counter()
    if p.returncode != 0:
        raise RuntimeError("reference CUDA backend failed:\n" + p.stdout[-2000:])
    text = p.stdout
    try:
        text += open(os.path.join(d, "log.dat")).read()
    except OSError:
        pass
    # "INFO: Total Running Time: X s, per step: Y ms" = the reference's own SimBackend timer (simulation loop only)
    m = re.search(r"Total Running Time:\s*([0-9.eE+-]+)\s*s", text)
    u = re.search(r"Lists updated\s*(\d+)\s*times", text)
    return (float(m.group(1)) if m else t1 - t0), ("SimBackend timer" if m else "wall clock"), (int(u.group(1)) if u else None)


def time_reference_cuda(top, conf, N, steps_a, steps_b, variants, T="300K", salt=0.5, dt=0.003, ext_forces=None, model_keys=None, binary=BIN):
    """variants: list of (use_edge, CUDA_sort_every).  Returns dict(best=..., runs=[...]) in particle-steps/s."""
    d = tempfile.mkdtemp(prefix="refcuda_")
    ext_path = None
    if ext_forces:
        ext_path = os.path.join(d, "forces.txt")
        write_forces_file(ext_path, ext_forces)
    runs = []
    for (use_edge, sort_every) in variants:
        try:
            ta, how, ua = _run(d, top, conf, steps_a, use_edge, sort_every, T, salt, dt, ext_path, model_keys, binary)
            tb, how, ub = _run(d, top, conf, steps_b, use_edge, sort_every, T, salt, dt, ext_path, model_keys, binary)
            val = N * (steps_b - steps_a) / max(tb - ta, 1e-9)
            runs.append(dict(use_edge=use_edge, CUDA_sort_every=sort_every, value=val, ms_per_md_step=1e3 * (tb - ta) / (steps_b - steps_a),
                             loop_s=[ta, tb], clock=how,
                             list_rebuild_every_md_steps=(steps_b - steps_a) / max(ub - ua, 1) if (ua is not None and ub is not None) else None))
        except Exception as e:  # pragma: no cover
            runs.append(dict(use_edge=use_edge, CUDA_sort_every=sort_every, value=None, error=str(e)[-400:]))
    ok = [r for r in runs if r.get("value")]
    best = max(ok, key=lambda r: r["value"]) if ok else None
    return dict(best=best, runs=runs, unit="particle-steps/s",
                method=f"stock CLI; difference of the reference's own 'Total Running Time' (SimBackend timer: simulation loop only) between a steps={steps_b} and a "
                       f"steps={steps_a} run, i.e. MD steps {steps_a}..{steps_b} after equilibration; " + ("timers WITHOUT device synchronisation (Timings.cpp -DNOCUDA)" if binary == BIN_NOSYNC else "timers on (a cudaDeviceSynchronize each), as users run it") + ", default threads_per_block")
