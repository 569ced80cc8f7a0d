import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from conftest import load_golden
import test_gpu_parity as T
g = load_golden("lattice8")
for nsteps in (1, 2, 5):
    a, b = T.make_sim(g), T.make_sim(g)
    for s in range(nsteps):
        a.ctx.set_step(s); a.ctx.first_step(); a.ctx.compute_forces(); a.ctx.second_step(); a.ctx.thermostat()
    b.run(nsteps)
    sa, sb = a.ctx.get_state(), b.ctx.get_state()
    print(os.environ.get("OXB_NO_GRAPHS"), nsteps, {k: float(np.abs(sa[k] - sb[k]).max()) for k in ("pos", "vel", "L", "a1")}, a.ctx.stats(), b.ctx.stats())
    a.close(); b.close()
