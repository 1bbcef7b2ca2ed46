"""The neuro-evolution loop of the reference's evolve/evolution.py:40-118 with everything
population-sized on the GPU(s): fitness evaluation (fly), best tracking, selection + crossover +
mutation and elitism.  One process per GPU, each rank flies and breeds its contiguous shard.

Multi-GPU parent fetch, two modes:
  * p2p (default when torch symmetric memory can be set up): every rank keeps ONLY its shard, in
    symmetric memory; the GA kernel (ld_ga_step_p2p_device) loads the tournament winners' rows
    straight out of the owning GPU's HBM over NVLink while it applies crossover and mutation.
    The only collective per generation is the all-gather of the 32-byte fitness records.
  * replicated: the genome matrix is replicated and the child shards are all-gathered (NCCL).

Differences from the reference, on purpose: random draws come from the library's counter-based
Philox stream, not numpy's global MT19937, and there is no jsonpickle dump — use
`best_genome()` / `learning_drones_b200.drone.drone.Drone.controller.set_genome`.
"""
import ctypes as C
import warnings

import numpy as np

from .. import _lib
from .._abi import GENOME_LEN, make_ga_params, make_params, make_peer_table
from ..distributed import allgather_records, shard_bounds
from ..population import fly_population_device, reduce_records_device


class GpuEvolution:
    def __init__(self, genomes, seed: int = 0, timesteps: int = 2500, mutation_rate: float = 0.05, p_tournament: float = 0.6,
                 p_crossover: float = 0.05, mutation_sigma: float = 0.3, device=None, p2p=None, local_shard: bool = False,
                 n_global: int = None, **fly_kw):
        """genomes: the whole (n, 314) population (same on every rank), or — local_shard=True — only this rank's
        contiguous shard of an n_global-drone population (big populations: no rank holds the whole matrix on the host)."""
        import torch
        import torch.distributed as dist

        self.torch = torch
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        g = torch.as_tensor(np.ascontiguousarray(genomes, dtype=np.float64)) if not torch.is_tensor(genomes) else genomes
        if g.dim() != 2 or g.shape[1] != GENOME_LEN:
            raise ValueError(f"genomes must have shape (n, {GENOME_LEN})")
        self.n = int(n_global) if local_shard else g.shape[0]
        if self.n % (2 * self.world) != 0:
            raise ValueError("population size must be a multiple of 2 x world size (pair slots, equal shards)")
        self.lo, self.hi = shard_bounds(self.n, self.rank, self.world)
        self.shard_rows = self.hi - self.lo
        if local_shard and g.shape[0] != self.shard_rows:
            raise ValueError(f"local shard has {g.shape[0]} rows, rank {self.rank} of {self.world} owns {self.shard_rows} of {self.n}")
        self._local_shard = bool(local_shard)
        slice_steps = fly_kw.pop("slice_steps", None)
        self.params = make_params(timesteps=timesteps, **fly_kw)
        # Scheduling only (ld_params.slice_steps).  Offspring populations mix a few long-lived drones with many
        # that crash early: fly them in doubling segments (boundaries H/4, 3H/4), also where the library's
        # automatic rule — which has to assume that every drone may survive — would not bother (+4 % generations/s
        # on populations with survivors).  While NO drone reaches the horizon (a young population: the whole of
        # BASELINE config 5's first generations) the segment queue only costs (-2.3 % generations/s, measured at
        # 131,072 drones): run() switches it off for the next block of generations whenever the last generation it
        # has read back had no survivor, and back on otherwise (_adapt_schedule).
        self._doubling = -max(125, timesteps // 4)
        self._adaptive_schedule = slice_steps is None
        self.params.slice_steps = self._doubling if slice_steps is None else int(slice_steps)
        self.seed, self.generation = int(seed), 0
        self.hyper = dict(p_tournament=p_tournament, p_crossover=p_crossover, mutation_rate=mutation_rate, mutation_sigma=mutation_sigma)
        # best-of-run tracking lives on the device (ld_track_best_device): {score, valid}, the genome, and one
        # row of statistics per generation, so that run() never has to wait for the GPU between generations
        self._best = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.global_best_genome = torch.empty(GENOME_LEN, dtype=torch.float64, device=self.device)
        self._stats = torch.empty(5, dtype=torch.int64, device=self.device)
        self._history = torch.zeros((0, 7), dtype=torch.int64, device=self.device)
        self.records = None  # fitness records of the generation flown last (whole population)
        # How the GA step reaches its parents — reported, never silent (bench.py prints it):
        #   "single"     one rank, the whole population in this GPU's memory
        #   "p2p"        shards in symmetric memory, parents loaded over NVLink inside the GA kernel
        #   "replicated" genome matrix replicated, child shards all-gathered with NCCL
        self.p2p, self.p2p_error = False, None
        if self.world > 1 and p2p is not False:
            try:
                self._setup_p2p(g)
                self.p2p = True
            except (RuntimeError, ImportError, AttributeError, NotImplementedError) as e:
                # torch symmetric memory cannot be set up on this system (no NVLink peer access, old torch)
                if p2p is True:
                    raise
                self.p2p_error = repr(e)
                warnings.warn(f"GpuEvolution: peer-memory GA step unavailable, using the replicated mode: {e!r}", RuntimeWarning)
        self.ga_mode = "single" if self.world == 1 else ("p2p" if self.p2p else "replicated")
        if not self.p2p:
            if local_shard and self.world > 1:
                self.population = torch.empty((self.n, GENOME_LEN), dtype=torch.float64, device=self.device)
                self.dist.all_gather_into_tensor(self.population, g.to(self.device, dtype=torch.float64).contiguous())
            else:
                self.population = g.to(self.device, dtype=torch.float64).contiguous().clone()
            self.next_population = torch.empty_like(self.population)

    # -- sharded storage in symmetric memory: peers read parent rows over NVLink --------------------
    def _setup_p2p(self, g):
        import torch.distributed._symmetric_memory as symm

        torch = self.torch
        group = self.dist.group.WORLD
        self._bufs, self._handles, self._tables = [], [], []
        for _ in range(2):  # parents / children, swapped every generation
            t = symm.empty(self.shard_rows, GENOME_LEN, dtype=torch.float64, device=self.device)
            h = symm.rendezvous(t, group.group_name)
            self._bufs.append(t)
            self._handles.append(h)
            self._tables.append(make_peer_table([int(p) for p in h.buffer_ptrs], self.shard_rows))
        self._cur = 0
        self._bufs[0].copy_((g if self._local_shard else g[self.lo:self.hi]).to(self.device, dtype=torch.float64))
        torch.cuda.synchronize(self.device)
        self.dist.barrier()

    @property
    def local_population(self):
        return self._bufs[self._cur] if self.p2p else self.population[self.lo:self.hi]

    @property
    def global_best_score(self):
        """Best reward seen so far (evolution.py:80-83); None before the first generation.  Synchronises."""
        score, valid = self._best.cpu().tolist()
        return score if valid else None

    # -- one generation: evolve/evolution.py:50-107, enqueued on the current stream without waiting ---
    def _enqueue_generation(self, history_row):
        torch = self.torch
        L = _lib.load()
        local_records, _ = fly_population_device(self.local_population, self.params)
        self.records = allgather_records(local_records, n_global=self.n) if self.dist else local_records
        reduce_records_device(self.records, self._stats)  # np.argmax(scores), evolution.py:77-78
        ga = make_ga_params(self.seed, self.generation, **self.hyper)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            # global best, :80-83 — the row is read from the current population (peer memory in p2p mode)
            _lib.check(L.ld_track_best_device(self._stats.data_ptr(), None if self.p2p else self.population.data_ptr(),
                                              C.byref(self._tables[self._cur]) if self.p2p else None, self.generation,
                                              self._best.data_ptr(), self.global_best_genome.data_ptr(), history_row.data_ptr(), stream))
            if self.p2p:
                children = self._bufs[1 - self._cur]
                _lib.check(L.ld_ga_step_p2p_device(C.byref(ga), C.byref(self._tables[self._cur]), self.records.data_ptr(), self.n,
                                                   children.data_ptr(), self.lo, self.hi, stream))
                if self.rank == 0:
                    children[0].copy_(self.global_best_genome)  # elitism, :105-107
                # peers may still be reading our parent shard; it is not written again before the next
                # generation's record all-gather, which every rank enters only after its own GA kernel
                self._cur = 1 - self._cur
            else:
                children = self.next_population[self.lo:self.hi]
                _lib.check(L.ld_ga_step_device(C.byref(ga), self.population.data_ptr(), self.records.data_ptr(), self.n,
                                               children.data_ptr(), self.lo, self.hi, stream))
                if self.dist:
                    self.dist.all_gather_into_tensor(self.next_population, children.clone())
                self.next_population[0].copy_(self.global_best_genome)  # elitism, :105-107
                self.population, self.next_population = self.next_population, self.population
        self.generation += 1

    @staticmethod
    def _history_dict(row):
        g, best, index, glob, steps, coins, crashed = row
        return dict(generation=g, best_score=best, best_index=index, global_best_score=glob, active_steps=steps, coins=coins,
                    crashed=crashed)

    def run(self, n_generations: int):
        """`n_generations` of the loop, enqueued back to back; one synchronisation at the end to read the
        per-generation statistics (a list of dicts like step()'s)."""
        torch = self.torch
        history = torch.zeros((n_generations, 7), dtype=torch.int64, device=self.device)
        for g in range(n_generations):
            self._enqueue_generation(history[g])
        self._history = torch.cat([self._history, history])
        rows = [self._history_dict(r) for r in history.cpu().tolist()]
        self._adapt_schedule(rows)
        return rows

    def _adapt_schedule(self, rows):
        """Episode segments for the next block of generations only if the last one had a survivor (scheduling only:
        results do not depend on it)."""
        if self._adaptive_schedule and rows:
            self.params.slice_steps = -1 if rows[-1]["crashed"] >= self.n else self._doubling

    def step(self):
        return self.run(1)[0]

    def best_genome(self) -> np.ndarray:
        return self.global_best_genome.cpu().numpy()

    def gather_population(self) -> np.ndarray:
        """Whole current population on the host (collective in multi-GPU runs)."""
        torch = self.torch
        if not self.dist:
            return self.population.cpu().numpy()
        if not self.p2p:
            return self.population.cpu().numpy()
        out = torch.empty((self.n, GENOME_LEN), dtype=torch.float64, device=self.device)
        self.dist.all_gather_into_tensor(out, self._bufs[self._cur].clone())
        return out.cpu().numpy()


class LocalShardedEvolution:
    """The sharded loop of GpuEvolution for `world` shards that all live on ONE GPU: every shard has its own
    buffers, the GA step and the best tracking address parents through an ld_peer_table exactly as they do over
    NVLink (ld_ga_step_p2p_device, ld_track_best_device(peers)).  Same results as GpuEvolution on one rank — which
    is what tests/test_gpu_ga.py checks, so the peer-memory code path is exercised on a single-GPU box too."""

    ga_mode = "p2p-local"

    def __init__(self, genomes, world: int, seed: int = 0, timesteps: int = 2500, mutation_rate: float = 0.05, p_tournament: float = 0.6,
                 p_crossover: float = 0.05, mutation_sigma: float = 0.3, device=None, **fly_kw):
        import torch

        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        g = torch.as_tensor(np.ascontiguousarray(genomes, dtype=np.float64)) if not torch.is_tensor(genomes) else genomes
        self.n, self.world = g.shape[0], int(world)
        if g.dim() != 2 or g.shape[1] != GENOME_LEN or self.n % (2 * self.world) != 0:
            raise ValueError(f"genomes must have shape (n, {GENOME_LEN}) with n a multiple of 2 x world")
        self.shard_rows = self.n // self.world
        slice_steps = fly_kw.pop("slice_steps", None)
        self.params = make_params(timesteps=timesteps, **fly_kw)
        self.params.slice_steps = -max(125, timesteps // 4) if slice_steps is None else int(slice_steps)
        self.seed, self.generation = int(seed), 0
        self.hyper = dict(p_tournament=p_tournament, p_crossover=p_crossover, mutation_rate=mutation_rate, mutation_sigma=mutation_sigma)
        rows = self.shard_rows
        self._bufs = [[g[r * rows:(r + 1) * rows].to(self.device, dtype=torch.float64).contiguous().clone() for r in range(self.world)],
                      [torch.empty((rows, GENOME_LEN), dtype=torch.float64, device=self.device) for _ in range(self.world)]]
        self._tables = [make_peer_table([t.data_ptr() for t in bufs], rows) for bufs in self._bufs]
        self._cur = 0
        self._best = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.global_best_genome = torch.empty(GENOME_LEN, dtype=torch.float64, device=self.device)
        self._stats = torch.empty(5, dtype=torch.int64, device=self.device)
        self.records = torch.empty((self.n, 4), dtype=torch.int64, device=self.device)

    def run(self, n_generations: int):
        torch = self.torch
        L = _lib.load()
        history = torch.zeros((n_generations, 7), dtype=torch.int64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rows = self.shard_rows
        with torch.cuda.device(self.device):
            for g in range(n_generations):
                cur, nxt, table = self._bufs[self._cur], self._bufs[1 - self._cur], self._tables[self._cur]
                for r in range(self.world):  # "all-gather": every shard's records land in population order
                    fly_population_device(cur[r], self.params, records=self.records[r * rows:(r + 1) * rows])
                reduce_records_device(self.records, self._stats)
                ga = make_ga_params(self.seed, self.generation, **self.hyper)
                _lib.check(L.ld_track_best_device(self._stats.data_ptr(), None, C.byref(table), self.generation, self._best.data_ptr(),
                                                  self.global_best_genome.data_ptr(), history[g].data_ptr(), stream))
                for r in range(self.world):
                    _lib.check(L.ld_ga_step_p2p_device(C.byref(ga), C.byref(table), self.records.data_ptr(), self.n, nxt[r].data_ptr(),
                                                       r * rows, (r + 1) * rows, stream))
                nxt[0][0].copy_(self.global_best_genome)  # elitism, evolution.py:105-107
                self._cur = 1 - self._cur
                self.generation += 1
        return [GpuEvolution._history_dict(r) for r in history.cpu().tolist()]

    def best_genome(self) -> np.ndarray:
        return self.global_best_genome.cpu().numpy()

    def gather_population(self) -> np.ndarray:
        return self.torch.cat(self._bufs[self._cur]).cpu().numpy()
