"""Diagnostic (not a test): where does the largest force error of the full-size C2 state come from?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
from oracle import oracle as O
from oxdna_b200 import lattice
from oxdna_b200.sim import Simulation, parse_temperature

sysm = lattice.duplex_lattice(2048, bp=20, spacing=10.0, seed=12345)
T = parse_temperature("300K")
v, L = lattice.maxwell_velocities(len(sysm["pos"]), T, 5)
inp = dict(backend="CUDA", interaction_type="DNA2", T="300K", salt_concentration=0.5, dt=0.003, verlet_skin=0.05, thermostat="brownian",
           newtonian_steps=103, diff_coeff=2.5, CUDA_sort_every=1, use_edge=int(sys.argv[1]) if len(sys.argv) > 1 else 1, seed=42)
sim = Simulation(inp, sysm, dict(box=sysm["box"], pos=sysm["pos"], a1=sysm["a1"], a3=sysm["a3"], vel=v, L=L))
sim.run(300)
st = sim.ctx.get_state()
P = O.dna2_params(T, 0.5)
ax = O.axes_from_a1a3(st["a1"], st["a3"])
pairs = O.verlet_pairs(st["pos"], sysm["n3"], sysm["n5"], sysm["box"], P.rcut + 0.1)
ref = O.forces(P, st["pos"], ax, sysm["btype"], sysm["n3"], sysm["n5"], sysm["box"], pairs)
sim.ctx.update_lists()
sim.ctx.compute_forces()
out = sim.ctx.get_forces()
err = np.linalg.norm(out["force"] - ref["force"], axis=1)
fn = np.linalg.norm(ref["force"], axis=1)
print("fmax", fn.max(), "max err", err.max(), "rel", err.max() / fn.max(), "median err", np.median(err), "99.9%", np.quantile(err, 0.999))
# backbone sites
back = st["pos"] + ax[:, 0:3] * (-0.34) + ax[:, 3:6] * 0.3408
for i in np.argsort(err)[::-1][:6]:
    msg = f"particle {i}: err {err[i]:.3e} |F| {fn[i]:.3f} dF {out['force'][i] - ref['force'][i]}"
    for nb, name in ((sysm["n3"][i], "n3"), (sysm["n5"][i], "n5")):
        if nb >= 0:
            d = np.linalg.norm(back[nb] - back[i])
            msg += f" | {name}={nb} rbb {d:.6f} x {d - 0.7564:.5f} err_nb {err[nb]:.2e}"
    print(msg)
# non-bonded-only and bonded-only comparison through the oracle: zero the pair list
ref_b = O.forces(P, st["pos"], ax, sysm["btype"], sysm["n3"], sysm["n5"], sysm["box"], pairs[:0])
i = int(np.argmax(err))
print("worst particle: bonded part of reference force", ref_b["force"][i], "full", ref["force"][i], "gpu", out["force"][i])
sim.close()

# ---- isolate the worst pair: the two particles alone (no bonds), same box and a smaller one
import numpy as np
a, b = (int(x) for x in np.argsort(err)[::-1][:2])
ids = [a, b]
none = np.full(2, -1, dtype=np.int32)
for Lbox in (130.0, 20.0):
    box = np.array([Lbox] * 3)
    topo = dict(btype=sysm["btype"][ids], n3=none, n5=none, strand=np.array([0, 1], dtype=np.int32))
    conf = dict(box=box, pos=st["pos"][ids], a1=st["a1"][ids], a3=st["a3"][ids])
    s2 = Simulation(dict(inp, thermostat="no"), topo, conf)
    o2 = s2.ctx.get_forces()
    ax2 = O.axes_from_a1a3(conf["a1"], conf["a3"])
    pr = O.verlet_pairs(conf["pos"], none, none, box, P.rcut + 0.1)
    r2 = O.forces(P, conf["pos"], ax2, topo["btype"], none, none, box, pr)
    print(f"pair ({a},{b}) alone, L = {Lbox}: |F| {np.linalg.norm(r2['force'][0]):.4f} err {np.linalg.norm(o2['force'] - r2['force'], axis=1)} terms {np.round(r2['eterms'], 5)}")
    print("   gpu split", np.round(s2.ctx.energy_split(), 5), " dr", np.linalg.norm(conf["pos"][0] - conf["pos"][1]))
    s2.close()
