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


class _Singles:
    """A list of independent replica objects (each with run(steps), system_energy(), update_temperature(T) -- one GPU context per replica,
    as in the reference's driver) presented through the batch interface below."""

    def __init__(self, replicas, concurrent):
        self.replicas = list(replicas)
        self.n_replicas = len(self.replicas)
        # local replicas advance CONCURRENTLY: one host thread per replica, each blocked inside the C ABI (ctypes drops the GIL)
        self.concurrent = concurrent and self.n_replicas > 1
        self._pool = None
        self._cur = [None] * self.n_replicas  # temperature each replica's Hamiltonian is currently set to (None: as constructed)

    def _map(self, fn, items):
        if not self.concurrent:
            return [fn(x) for x in items]
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.n_replicas)
        return list(self._pool.map(fn, items))

    def run(self, steps):
        self._map(lambda rep: rep.run(steps), self.replicas)

    def energies(self):
        return np.array(self._map(lambda rep: rep.system_energy(), self.replicas), dtype=np.float64)

    def _set(self, k, T):
        if self._cur[k] is None or self._cur[k] != T:
            self.replicas[k].update_temperature(T)
            self._cur[k] = T

    def energies_at(self, T, mask):
        out = np.zeros(self.n_replicas)

        def one(k):
            if mask[k]:
                self._set(k, T[k])
                out[k] = self.replicas[k].system_energy()

        self._map(one, range(self.n_replicas))
        return out

    def set_temperatures(self, T, first=False):
        for k in range(self.n_replicas):
            if first:
                self._cur[k] = T[k]  # the replicas were constructed at their ladder temperature
            else:
                self._set(k, T[k])


class ReplicaExchange:
    """replicas: the LOCAL replicas of this rank, either
      * a list of BATCHES (objects with n_replicas, run(steps), energies(), energies_at(T, mask), set_temperatures(T)): oxdna_b200.sim.ReplicaBatch
        holds R replicas in ONE GPU context and advances them with one launch per kernel -- the B200-native layout; or
      * a list of single replica objects with run(steps), system_energy(), update_temperature(T) (one context each, the reference's layout).
    temperatures: the global ladder (simulation units), len = world_size * number of local replicas.
    Global replica id g = rank * n_local + k; initially replica g sits on ladder position g."""

    def __init__(self, replicas, temperatures, comm=None, seed=0, concurrent=True):
        self.comm = comm or LocalComm()
        replicas = list(replicas)
        if replicas and hasattr(replicas[0], "n_replicas"):
            self.groups = replicas
        else:
            if concurrent and len(replicas) > 1:
                # one driver thread per local replica on every rank of the node: once they outnumber the cores, spinning waits starve
                # each other, so the threads sleep on blocking-sync events instead
                import os
                n_threads = len(replicas) * int(os.environ.get("LOCAL_WORLD_SIZE", self.comm.world_size))
                if n_threads > (os.cpu_count() or 1):
                    for rep in replicas:
                        ctx = getattr(rep, "ctx", None)
                        if ctx is not None and hasattr(ctx, "set_host_wait"):
                            ctx.set_host_wait(True)
            self.groups = [_Singles(replicas, concurrent)]
        self.replicas = replicas
        self.nl = sum(g.n_replicas for g in self.groups)
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
        # host wall-clock per phase on this rank (seconds): MD, energy evaluations of the exchange, waiting in / for the collective
        self.timers = dict(md=0.0, energy=0.0, comm=0.0, update=0.0)
        off = 0
        for g in self.groups:
            g.set_temperatures(self.T[self.comm.rank * self.nl + off + np.arange(g.n_replicas)], first=True)
            off += g.n_replicas

    def _gid(self, k):
        return self.comm.rank * self.nl + k

    def exchange(self):
        """One exchange attempt (no MD).  Returns the list of accepted ladder pairs."""
        import time
        pairs = attempted_pairs(self.round, self.R)
        at = {int(self.location[g]): g for g in range(self.R)}  # ladder position -> replica
        uppers = {b for (_, b) in pairs}
        pos = np.array([int(self.location[self._gid(k)]) for k in range(self.nl)])
        is_upper = np.array([p in uppers for p in pos])
        local = np.zeros((self.nl, 2))
        t0 = time.perf_counter()
        off = 0
        for g in self.groups:
            sl = slice(off, off + g.n_replicas)
            local[sl, 0] = g.energies()
            if is_upper[sl].any():
                # upper member of an attempted pair: energy with the partner's (lower) temperature Hamiltonian
                T_alt = np.where(is_upper[sl], self.T[np.maximum(pos[sl] - 1, 0)], self.T[pos[sl]])
                local[sl, 1] = g.energies_at(T_alt, is_upper[sl])
            off += g.n_replicas
        t1 = time.perf_counter()
        E = self.comm.all_gather(local.reshape(-1)).reshape(self.R, 2)
        t2 = time.perf_counter()
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
        off = 0
        for g in self.groups:
            gids = self.comm.rank * self.nl + off + np.arange(g.n_replicas)
            g.set_temperatures(self.T[self.location[gids]])
            off += g.n_replicas
        t3 = time.perf_counter()
        self.timers["energy"] += t1 - t0
        self.timers["comm"] += t2 - t1
        self.timers["update"] += t3 - t2
        self.history.append(self.location.copy())
        self.round += 1
        return accepted

    def advance(self, steps):
        """`steps` MD steps of every local replica"""
        import time
        t0 = time.perf_counter()
        if len(self.groups) == 1:
            self.groups[0].run(steps)
        else:
            # a few large batches per GPU (a batch is capped by the 22-bit particle index): one host thread each, so that one batch's
            # host synchronisations are hidden behind the other's kernels
            if getattr(self, "_gpool", None) is None:
                from concurrent.futures import ThreadPoolExecutor
                self._gpool = ThreadPoolExecutor(max_workers=len(self.groups))
            list(self._gpool.map(lambda g: g.run(steps), self.groups))
        self.timers["md"] += time.perf_counter() - t0

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
