"""Diagnostic: does the unmodified reference CUDA backend rebuild its Verlet lists on this box, and do particles move?"""
import sys, os, subprocess, re
import numpy as np
sys.path.insert(0, os.getcwd())
import bench
from oracle import ref_cuda_bench as R
from oxdna_b200 import io as oio
sysm, desc = bench.workload("small")
d = "/tmp/reflog"; os.makedirs(d, exist_ok=True)
top, conf = bench.write_case(sysm, bench.T_STR, d)
c0 = oio.read_conf(conf)
for backend in ("CUDA",):
    t = R.TEMPLATE.format(salt=0.5, T="300K", dt=0.003, steps=12, sort_every=0, use_edge=1, top=top, conf=conf, d=d, ext=0, extfile="")
    t = t.replace("backend = CUDA", "backend = " + backend).replace("print_energy_every = 100000000", "print_energy_every = 1").replace("CUDA_avoid_cpu_calculations = 1", "CUDA_avoid_cpu_calculations = 0").replace("use_edge = 1", "use_edge = " + os.environ.get("UE", "1")).replace("thermostat = brownian", "thermostat = " + os.environ.get("TH", "brownian"))
    open(os.path.join(d, "input"), "w").write(t)
    p = subprocess.run([R.BIN, "input"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = p.stdout + open(os.path.join(d, "log.dat")).read()
    print(backend, p.returncode, p.stdout[-600:] if p.returncode else "", sorted(os.listdir(d)), re.findall(r"Lists updated.*|Total Running.*|\*\*\*> Lists.*", log))
    if not os.path.exists(os.path.join(d, "last_conf.dat")):
        print(log[-1500:]); continue
    c1 = oio.read_conf(os.path.join(d, "last_conf.dat"))
    dr = np.linalg.norm(c1["pos"] - c0["pos"], axis=1)
    print("  displacement after 3000 steps: mean %.4f max %.4f;  |v| rms %.4f" % (dr.mean(), dr.max(), np.sqrt((c1["vel"] ** 2).sum(1).mean())))
    print("  energy.dat:", open(os.path.join(d, "energy.dat")).read().strip().replace("\n", " | "))
