"""Prints the force/torque/energy error of the CUDA path against the committed reference fixtures (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from test_gpu_parity import make_sim
for case in ["force_field_dna/ref_dna2_nomesh", "lattice8", "lattice27_dense"]:
    g = load_golden(case)
    for ue in (0, 1):
        sim = make_sim(g, use_edge=ue, CUDA_sort_every=1)
        o = sim.ctx.get_forces()
        fmax = np.linalg.norm(g["force"], axis=1).max(); tmax = np.linalg.norm(g["torque_lab"], axis=1).max()
        print(f"{case:34s} edge={ue} dF/Fmax={np.linalg.norm(o['force']-g['force'],axis=1).max()/fmax:.2e} "
              f"dT/Tmax={np.linalg.norm(o['torque_lab']-g['torque_lab'],axis=1).max()/tmax:.2e} dU/U={abs(o['U']-float(g['U']))/abs(float(g['U'])):.2e}")
        sim.close()
