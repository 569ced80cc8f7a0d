"""GPU tests of the plugin seam of the C ABI (SURVEY 8 a13 / b): oxb_device_views and oxb_set_force_callback.

The reference hands its raw device arrays (d_poss float4, d_orientations GPU_quat, matrix_neighs column-major, number_neighs) to
whatever CUDABaseInteraction its factory returned -- for an unknown interaction_type the class PluginManager finds in CUDA<type>.so
(src/CUDA/Interactions/CUDAInteractionFactory.cu:44-51, src/CUDA/Interactions/CUDABaseInteraction.h:60).  Here a toy third-party
interaction -- a soft repulsion k (rc - r)^2 / 2 between every listed pair, written with torch ops on the context's own stream -- reads
those views zero-copy and writes the force accumulators; the trajectory is compared with a plain numpy O(N^2) velocity-Verlet run of
the same potential.  (The C++ side of the seam, make_CUDA<type> through the reference's PluginManager, is tested in test_dropin.py.)"""
import numpy as np
import pytest

from conftest import load_golden, pair_set
from oracle import oracle as O
from oxdna_b200 import capi

pytestmark = pytest.mark.gpu

K_TOY, RC_TOY = 3.0, 1.4


class _Dev:
    """a raw device pointer as a __cuda_array_interface__ object (zero-copy torch view)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(int(ptr), False), version=3, strides=None)


def _view(ptr, shape, typestr):
    import torch
    return torch.as_tensor(_Dev(ptr, shape, typestr), device="cuda")


def toy_force_pass(ctx, calls):
    """the third-party force pass: everything it knows comes from the oxb_force_views block"""
    import torch

    def fn(v):
        N, rows = v.N, ctx.stats()["max_neigh"]
        with torch.cuda.stream(torch.cuda.ExternalStream(v.stream)):
            pos = _view(v.poss, (N, 4), "<f4")[:, :3]
            M = _view(v.matrix_neighs, (rows, v.stride), "<i4")  # column-major: M[k, i] = k-th neighbour of slot i
            nn = _view(v.number_neighs, (N,), "<i4")
            F = _view(v.forces, (N, 4), "<f4")
            box = torch.tensor(list(v.box), dtype=torch.float32, device="cuda")
            kmax = int(nn.max().item())
            valid = torch.arange(kmax, device="cuda")[:, None] < nn[None, :]
            j = torch.where(valid, M[:kmax], torch.zeros_like(M[:kmax])).long()
            d = pos[j] - pos[None, :, :]
            d = d - box * torch.round(d / box)
            r = d.norm(dim=2)
            on = valid & (r < RC_TOY)
            mag = torch.where(on, K_TOY * (RC_TOY - r) / r.clamp_min(1e-6), torch.zeros_like(r))
            F[:, :3] += -(mag[:, :, None] * d).sum(dim=0)
            F[:, 3] += torch.where(on, 0.5 * K_TOY * (RC_TOY - r) ** 2, torch.zeros_like(r)).sum(dim=0)
        calls.append(v.step)
        return 0
    return fn


def toy_reference(pos, vel, n3, n5, box, dt, steps):
    """numpy O(N^2) velocity Verlet of the same potential (bonded neighbours are not in the Verlet lists: excluded)"""
    N = len(pos)
    excl = np.eye(N, dtype=bool)
    for i in range(N):
        for b in (n3[i], n5[i]):
            if b >= 0:
                excl[i, b] = True

    def forces(p):
        d = p[None, :, :] - p[:, None, :]
        d -= box * np.round(d / box)
        r = np.linalg.norm(d, axis=2)
        on = (r < RC_TOY) & ~excl
        mag = np.where(on, K_TOY * (RC_TOY - r) / np.where(r > 0, r, 1.0), 0.0)
        return -(mag[:, :, None] * d).sum(axis=1), 0.5 * np.where(on, 0.5 * K_TOY * (RC_TOY - r) ** 2, 0.0).sum()

    p, v = pos.copy(), vel.copy()
    f, _ = forces(p)
    for _ in range(steps):
        v += 0.5 * dt * f
        p += dt * v
        f, U = forces(p)
        v += 0.5 * dt * f
    return p, v, U


@pytest.mark.parametrize("sort_every", [0, 1])
def test_toy_plugin_reads_the_views_and_drives_the_step(sort_every):
    g = load_golden("lattice8")
    N = len(g["pos"])
    ctx = capi.Context(N)
    calls = []
    try:
        ctx.set_box(g["box"])
        ctx.set_topology(g["btype"], g["n3"], g["n5"], g["strand"])
        ctx.set_force_callback(toy_force_pass(ctx, calls), RC_TOY)
        ctx.set_lists(verlet_skin=0.05, use_edge=False, sort_every=sort_every)
        ctx.set_dt(0.003)
        ctx.set_thermostat(capi.THERMOSTAT_NONE if hasattr(capi, "THERMOSTAT_NONE") else 0)
        ctx.set_state(g["pos"], g["a1"], g["a3"], g["vel"], g["L"])
        # the Verlet matrix the plugin sees is the reference's pair set for ITS cutoff
        assert pair_set(ctx.get_pairs()) == pair_set(O.verlet_pairs(g["pos"], g["n3"], g["n5"], g["box"], RC_TOY + 2 * 0.05))
        steps = 300
        ctx.run(steps)
        st = ctx.get_state()
        U, _ = ctx.energy()
        p_ref, v_ref, U_ref = toy_reference(g["pos"], g["vel"], g["n3"], g["n5"], g["box"], 0.003, steps)
        assert len(calls) >= steps  # one force pass per step went through the callback
        assert np.abs(st["pos"] - p_ref).max() < 2e-5 and np.abs(st["vel"] - v_ref).max() < 2e-5
        assert abs(U - U_ref) < 1e-5 * max(1.0, abs(U_ref))
        # no torque from this potential: the angular momenta are untouched
        assert np.abs(st["L"] - g["L"]).max() < 1e-12
        if sort_every:
            assert ctx.stats()["n_sorts"] >= 1
    finally:
        ctx.close()


def test_device_views_expose_the_reference_layouts():
    """oxb_device_views on an edge-pipeline context (whose own builds keep half the matrix): both directions of every pair are there"""
    import torch
    from oxdna_b200.sim import Simulation
    g = load_golden("lattice8")
    inp = dict(backend="CUDA", interaction_type="DNA2", T=str(g["T"]), salt_concentration=float(g["salt"]), dt=0.003, verlet_skin=0.05,
               thermostat="no", CUDA_sort_every=1, use_edge=1, seed=11)
    sim = Simulation(inp, dict(btype=g["btype"], n3=g["n3"], n5=g["n5"], strand=g["strand"]),
                     dict(box=g["box"], pos=g["pos"], a1=g["a1"], a3=g["a3"], vel=g["vel"], L=g["L"]))
    try:
        sim.ctx.compute_forces()
        N = len(g["pos"])
        v = sim.ctx.device_views()
        rows = sim.ctx.stats()["max_neigh"]
        pos = _view(v["poss"], (N, 4), "<f4").cpu().numpy()
        quat = _view(v["orientations"], (N, 4), "<f4").cpu().numpy()
        M = _view(v["matrix_neighs"], (rows, N), "<i4").cpu().numpy()
        nn = _view(v["number_neighs"], (N,), "<i4").cpu().numpy()
        word = pos[:, 3].copy().view(np.int32)
        orig = word & 0x003FFFFF  # MD_CUDABackend.cu:243-254
        assert sorted(orig.tolist()) == list(range(N))
        assert np.array_equal(word >> 22, g["btype"][orig])
        assert np.abs(pos[:, :3] - g["pos"][orig]).max() < 1e-5
        # quaternion -> a1 (src/CUDA/cuda_utils/CUDA_lr_common.cuh:40-61)
        x, y, z, w = quat.T
        a1 = np.stack([x * x - y * y - z * z + w * w, 2 * (x * y + z * w), 2 * (x * z - y * w)], axis=1)
        assert np.abs(a1 - g["a1"][orig]).max() < 1e-6
        both = set()
        for s in range(N):
            for k in range(nn[s]):
                both.add((int(orig[s]), int(orig[M[k, s]])))
        ref = pair_set(g["pairs"])
        assert both == ref | {(b, a) for a, b in ref}
        torch.cuda.synchronize()
        # the forces of the edge pipeline are unaffected by the switch to full builds
        fmax = np.linalg.norm(g["force"], axis=1).max()
        assert np.linalg.norm(sim.ctx.get_forces()["force"] - g["force"], axis=1).max() <= 1e-5 * fmax
    finally:
        sim.close()
