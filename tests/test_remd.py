"""Replica-exchange logic (oxdna_b200/remd.py) on the CPU: exchange rule of the reference's examples/OXPY_REMD/remd.py
(lines 23-37, 104-147) and the multi-process path over torch.distributed with the gloo backend, world_size = 2.
Replicas are stand-ins with an analytic temperature-dependent energy (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oxdna_b200.remd import LocalComm, ReplicaExchange, acceptance, attempted_pairs, geometric_ladder  # noqa: E402


class FakeReplica:
    """U(T) = e0 + c * T_hamiltonian: mimics a temperature-dependent Hamiltonian (stacking strength depends on T)."""

    def __init__(self, e0, T):
        self.e0, self.T, self.ran = e0, T, 0
        self.history = []

    def run(self, steps):
        self.ran += steps

    def system_energy(self):
        return self.e0 + 3.0 * self.T

    def update_temperature(self, T):
        self.T = T
        self.history.append(T)


def test_pairs_alternate_like_the_reference():
    # remd.py:23-37: even rounds attempt (1,2),(3,4)..., odd rounds (0,1),(2,3)...
    assert attempted_pairs(0, 6) == [(1, 2), (3, 4)]
    assert attempted_pairs(1, 6) == [(0, 1), (2, 3), (4, 5)]
    assert attempted_pairs(0, 2) == []
    assert attempted_pairs(1, 2) == [(0, 1)]


def test_acceptance_rule():
    # min(1, exp((1/Ta - 1/Tb) (Ea - Eb|Ta)))
    assert acceptance(0.10, 0.11, -10.0, -12.0) == 1.0
    x = (1 / 0.10 - 1 / 0.11) * (-12.0 + 10.0)
    assert abs(acceptance(0.10, 0.11, -12.0, -10.0) - np.exp(x)) < 1e-15


def test_geometric_ladder():
    t = geometric_ladder(290.0, 350.0, 64)
    assert abs(t[0] - 290.0) < 1e-12 and abs(t[-1] - 350.0) < 1e-12
    assert np.allclose(t[1:] / t[:-1], (350.0 / 290.0) ** (1 / 63))


def run_exchanges(comm, n_local, rounds, seed=7):
    R = n_local * comm.world_size
    T = geometric_ladder(0.09, 0.12, R)
    e0 = -np.arange(R, dtype=float)  # hotter ladder positions start with lower energy: swaps are favourable
    reps = [FakeReplica(e0[comm.rank * n_local + k], T[comm.rank * n_local + k]) for k in range(n_local)]
    rx = ReplicaExchange(reps, T, comm, seed=seed)
    acc = []
    for _ in range(rounds):
        acc.append(rx.exchange())
    return rx, reps, acc


def test_single_process_exchange_moves_temperatures_not_configurations():
    rx, reps, acc = run_exchanges(LocalComm(), 4, 6)
    # location stays a permutation of the ladder, every replica's temperature equals its ladder position's
    assert sorted(rx.location.tolist()) == [0, 1, 2, 3]
    for g, rep in enumerate(reps):
        assert abs(rep.T - rx.T[rx.location[g]]) < 1e-15
    assert sum(len(a) for a in acc) > 0
    # the one-sided rule evaluated the upper member at its partner's temperature and then restored / swapped it
    assert any(len(r.history) >= 2 for r in reps)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oxdna_b200.remd import TorchComm
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rx, reps, acc = run_exchanges(TorchComm(), 2, 6)
        q.put((rank, rx.location.tolist(), [r.T for r in reps], acc, rx.rates().tolist()))
    finally:
        dist.destroy_process_group()


def test_two_process_gloo_matches_single_process():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, ref_reps, ref_acc = run_exchanges(LocalComm(), 4, 6)
    # both ranks took identical decisions, equal to the single-process run with the same seed
    assert out[0][1] == out[1][1] == ref.location.tolist()
    assert out[0][3] == out[1][3] == ref_acc
    temps = out[0][2] + out[1][2]
    assert np.allclose(temps, [r.T for r in ref_reps])
    assert np.allclose(out[0][4], ref.rates())
