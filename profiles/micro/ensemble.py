"""micro-benchmark: R replicas of C2 on ONE GPU, sequential vs concurrent host threads (round 1, C5 design check)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from oxdna_b200 import lattice
from oxdna_b200.remd import ReplicaExchange, geometric_ladder
from oxdna_b200.sim import Simulation, parse_temperature

class A: pass
args = A(); args.sort_every = 1; args.use_edge = 1; args.workload = "c2"
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
md = int(sys.argv[2]) if len(sys.argv) > 2 else 500
sysm, desc = bench.workload("c2")
N = len(sysm["pos"])
ladder = geometric_ladder(290.0, 350.0, R)
sims = []
for r in range(R):
    T = f"{ladder[r]:.6f}K"
    v, L = lattice.maxwell_velocities(N, parse_temperature(T), 5 + r)
    sims.append(Simulation(bench.base_input(args, T), sysm, dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L)))
for conc in (False, True):
    re = ReplicaExchange(sims, ladder * 0.1 / 300.0, None, seed=1, concurrent=conc)
    re.advance(3000 if not conc else 500)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        re.advance(md); re.exchange()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"R={R} concurrent={conc}: {R * N * md * 3 / (t1 - t0):.4g} particle-steps/s  ({(t1 - t0) / 3 * 1e3:.1f} ms per round of {md} steps)", flush=True)
