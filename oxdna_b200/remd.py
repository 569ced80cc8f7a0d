"""Replica-exchange MD (parallel tempering in temperature) across GPUs.

B200-native restatement of the reference's examples/OXPY_REMD/remd.py (mpi4py, one replica per process):
 * R replicas are block-distributed over the G ranks of a torch.distributed group (one process per GPU, backend
   "nccl"; "gloo" in the CPU tests), R/G replicas per GPU;
 * per exchange round the only traffic is ONE all_gather of 2 doubles per replica (energy at the replica's own
   temperature, and -- for the upper member of each attempted pair -- at its partner's temperature);
 * swap decisions are computed redundantly on every rank from a shared counter-based random stream, so no second
   message is needed; temperatures move, configurations stay (as in the reference).

Exchange rule (remd.py:23-37,104-147): in round i the temperature-ladder pairs (a, a+1) with a % 2 == (i % 2 == 0)
are attempted.  The "responsible" replica (ladder position a) uses its energy E_a at its own temperature; the partner
(position a+1) switches its Hamiltonian to T_a, evaluates E_{a+1}|T_a, and the swap is accepted with probability
min(1, exp((1/T_a - 1/T_{a+1}) (E_a - E_{a+1}|T_a))).  On acceptance the two replicas exchange temperatures.
"""
import numpy as np


class LocalComm:
    """Single-process stand-in for a torch.distributed group."""
    rank, world_size = 0, 1

    def all_gather(self, local):
        return np.asarray(local, dtype=np.float64).copy()


class TorchComm:
    """all_gather of a float64 vector over torch.distributed (NCCL on the GPU, gloo on the CPU)."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
        self.device = device

    def all_gather(self, local):
        import torch
        t = torch.as_tensor(np.asarray(local, dtype=np.float64))
        if self.device is not None:
            t = t.to(self.device)
        out = torch.empty(self.world_size * t.numel(), dtype=torch.float64, device=t.device)
        self.dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy()


def attempted_pairs(round_index, n_temps):
    """Ladder positions (a, a+1) attempted in this round -- remd.py:23-37."""
    odd_pairs = 1 if (round_index % 2) == 0 else 0
    return [(a, a + 1) for a in range(n_temps - 1) if (a % 2) == odd_pairs]


def acceptance(T_a, T_b, E_a, E_b_at_Ta):
    x = (1.0 / T_a - 1.0 / T_b) * (E_a - E_b_at_Ta)
    return 1.0 if x >= 0 else float(np.exp(x))


class ReplicaExchange:
    """replicas: the LOCAL replica objects of this rank, each with run(steps), system_energy(), update_temperature(T).
    temperatures: the global ladder (simulation units), len = world_size * len(replicas).
    Global replica id g = rank * n_local + k; initially replica g sits on ladder position g."""

    def __init__(self, replicas, temperatures, comm=None, seed=0, concurrent=True):
        self.comm = comm or LocalComm()
        # local replicas advance CONCURRENTLY: one host thread per replica, each blocked inside the C ABI (ctypes drops the
        # GIL), each context on its own CUDA streams -- an 81,920-nt replica alone cannot fill 148 SMs
        self.concurrent = concurrent and len(replicas) > 1
        self._pool = None
        if self.concurrent:
            # one driver thread per local replica on every rank of the node: once they outnumber the cores, spinning waits starve
            # each other (64 replicas on 8 GPUs of a 16-core host), so the threads sleep on blocking-sync events instead
            import os
            n_threads = len(replicas) * int(os.environ.get("LOCAL_WORLD_SIZE", self.comm.world_size))
            if n_threads > (os.cpu_count() or 1):
                for rep in replicas:
                    ctx = getattr(rep, "ctx", None)
                    if ctx is not None and hasattr(ctx, "set_host_wait"):
                        ctx.set_host_wait(True)
        self.replicas = list(replicas)
        self.nl = len(self.replicas)
        self.T = np.asarray(temperatures, dtype=np.float64)
        self.R = len(self.T)
        if self.R != self.nl * self.comm.world_size:
            raise ValueError(f"The number of temperatures ({self.R}) should match the number of replicas ({self.nl * self.comm.world_size})")
        self.location = np.arange(self.R)  # location[g] = ladder position of replica g (identical on all ranks)
        self.seed = seed
        self.round = 0
        self.tries = np.zeros(self.R)
        self.accepts = np.zeros(self.R)
        self.history = []

    def _gid(self, k):
        return self.comm.rank * self.nl + k

    def exchange(self):
        """One exchange attempt (no MD).  Returns the list of accepted ladder pairs."""
        pairs = attempted_pairs(self.round, self.R)
        at = {int(self.location[g]): g for g in range(self.R)}  # ladder position -> replica
        local = np.zeros((self.nl, 2))

        def energies(k):
            rep = self.replicas[k]
            pos = int(self.location[self._gid(k)])
            local[k, 0] = rep.system_energy()
            if any(pos == b for (_, b) in pairs):
                # upper member of an attempted pair: energy with the partner's (lower) temperature Hamiltonian
                rep.update_temperature(self.T[pos - 1])
                local[k, 1] = rep.system_energy()
                rep.update_temperature(self.T[pos])

        self._map(energies, range(self.nl))
        E = self.comm.all_gather(local.reshape(-1)).reshape(self.R, 2)
        rng = np.random.default_rng([self.seed, self.round])
        u = rng.random(self.R)
        accepted = []
        for (a, b) in pairs:
            ga, gb = at[a], at[b]
            p = acceptance(self.T[a], self.T[b], E[ga, 0], E[gb, 1])
            self.tries[ga] += 1
            self.tries[gb] += 1
            if u[a] < p:
                accepted.append((a, b))
                self.location[ga], self.location[gb] = b, a
                self.accepts[ga] += 1
                self.accepts[gb] += 1
        for k, rep in enumerate(self.replicas):
            g = self._gid(k)
            if any(g in (at[a], at[b]) for (a, b) in accepted):
                rep.update_temperature(self.T[int(self.location[g])])
        self.history.append(self.location.copy())
        self.round += 1
        return accepted

    def _map(self, fn, items):
        if not self.concurrent:
            return [fn(x) for x in items]
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.nl)
        return list(self._pool.map(fn, items))

    def advance(self, steps):
        """`steps` MD steps of every local replica"""
        self._map(lambda rep: rep.run(steps), self.replicas)

    def run(self, rounds, pt_move_every):
        for _ in range(rounds):
            self.advance(pt_move_every)
            self.exchange()

    def rates(self):
        return self.accepts / np.maximum(self.tries, 1)


def geometric_ladder(t_lo, t_hi, n):
    if n == 1:
        return np.array([t_lo])
    return t_lo * (t_hi / t_lo) ** (np.arange(n) / (n - 1))
