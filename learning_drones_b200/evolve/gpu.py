"""The neuro-evolution loop of the reference's evolve/evolution.py:40-118 with everything
population-sized on the GPU(s): fitness evaluation (fly), best tracking, selection + crossover +
mutation (ld_ga_step_device) and elitism.  One process per GPU; the genome matrix is replicated,
each rank flies and breeds its contiguous shard and the shards are all-gathered (NCCL).

Differences from the reference, on purpose: random draws come from the library's counter-based
Philox stream, not numpy's global MT19937, and there is no jsonpickle dump — use
`best_genome()` / `learning_drones_b200.drone.drone.Drone.controller.set_genome`.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .._abi import GENOME_LEN, make_ga_params, make_params
from ..distributed import allgather_records, shard_bounds
from ..population import fly_population_device, reduce_records


class GpuEvolution:
    def __init__(self, genomes, seed: int = 0, timesteps: int = 2500, mutation_rate: float = 0.05, p_tournament: float = 0.6,
                 p_crossover: float = 0.05, mutation_sigma: float = 0.3, device=None, **fly_kw):
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
        self.n = g.shape[0]
        if self.n % (2 * self.world) != 0:
            raise ValueError("population size must be a multiple of 2 x world size (pair slots, equal shards)")
        self.population = g.to(self.device, dtype=torch.float64).contiguous().clone()
        self.next_population = torch.empty_like(self.population)
        self.lo, self.hi = shard_bounds(self.n, self.rank, self.world)
        self.params = make_params(timesteps=timesteps, **fly_kw)
        self.seed, self.generation = int(seed), 0
        self.hyper = dict(p_tournament=p_tournament, p_crossover=p_crossover, mutation_rate=mutation_rate, mutation_sigma=mutation_sigma)
        self.global_best_score = None
        self.global_best_genome = torch.empty(GENOME_LEN, dtype=torch.float64, device=self.device)
        self.records = None  # fitness records of the generation flown last (whole population)

    # -- one generation: evolve/evolution.py:50-107 ------------------------------------------------
    def step(self):
        torch = self.torch
        L = _lib.load()
        local = self.population[self.lo:self.hi]
        local_records, _ = fly_population_device(local, self.params)
        self.records = allgather_records(local_records, n_global=self.n) if self.dist else local_records
        stats = reduce_records(self.records)  # np.argmax(scores), evolution.py:77-78
        best_score, best_index = stats["best_reward"], stats["best_index"]
        if self.global_best_score is None or best_score > self.global_best_score:  # :80-83
            self.global_best_score = best_score
            self.global_best_genome.copy_(self.population[best_index])
        ga = make_ga_params(self.seed, self.generation, **self.hyper)
        children = self.next_population[self.lo:self.hi]
        with torch.cuda.device(self.device):
            _lib.check(L.ld_ga_step_device(C.byref(ga), self.population.data_ptr(), self.records.data_ptr(), self.n,
                                           children.data_ptr(), self.lo, self.hi, torch.cuda.current_stream(self.device).cuda_stream))
        if self.dist:
            self.dist.all_gather_into_tensor(self.next_population, children.clone())
        self.next_population[0].copy_(self.global_best_genome)  # elitism, :105-107
        self.population, self.next_population = self.next_population, self.population
        self.generation += 1
        return dict(generation=self.generation - 1, best_score=best_score, best_index=best_index,
                    global_best_score=self.global_best_score, active_steps=stats["active_steps"], coins=stats["coins"],
                    crashed=stats["crashed"])

    def run(self, n_generations: int):
        return [self.step() for _ in range(n_generations)]

    def best_genome(self) -> np.ndarray:
        return self.global_best_genome.cpu().numpy()
