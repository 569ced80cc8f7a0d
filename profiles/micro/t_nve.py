import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from conftest import load_golden
import test_gpu_parity as T
g = load_golden("lattice8")
for ue, se in ((0,0),(1,0),(1,1),(0,1)):
    sim = T.make_sim(g, use_edge=ue, CUDA_sort_every=se)
    n = int(g["nve_steps"])
    sim.run(n)
    st = sim.ctx.get_state()
    U, K = sim.ctx.energy()
    E0 = float(g["U"]) + 0.5 * (np.sum(g["vel"] ** 2) + np.sum(g["L"] ** 2))
    print(os.environ.get("OXB_NO_GRAPHS"), ue, se, "n", n, "dpos %.2e dvel %.2e da1 %.2e dE %.2e" % (np.abs(st["pos"] - g["pos1"]).max(), np.abs(st["vel"] - g["vel1"]).max(), np.abs(st["a1"] - g["a11"]).max(), abs(U+K-E0)/abs(E0)), sim.ctx.stats())
    sim.close()
