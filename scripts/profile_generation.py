#!/usr/bin/env python
"""Where does a generation of the GPU evolution loop spend its time? (dev tool, single GPU)"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from learning_drones_b200 import _lib
from learning_drones_b200._abi import make_ga_params
from learning_drones_b200.evolve.gpu import GpuEvolution
from learning_drones_b200.population import fly_population_device, reduce_records

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
rng = np.random.default_rng(0)
genomes = np.concatenate([rng.normal(0.0, 1.0 / np.sqrt(r), size=(n, r * c)) for r, c in ((10, 16), (10, 11), (4, 11))], axis=1)
ev = GpuEvolution(genomes, seed=0, timesteps=2000)
ev.run(20)  # evolve a bit so the population looks like mid-run offspring
torch.cuda.synchronize()
L = _lib.load()
ev_ = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
acc = np.zeros(4)
wall0 = time.perf_counter()
G = 30
for _ in range(G):
    ev_[0].record()
    rec, _ = fly_population_device(ev.local_population, ev.params)
    ev_[1].record()
    stats = reduce_records(rec)  # includes the D2H of 5 numbers (sync)
    ev_[2].record()
    ga = make_ga_params(ev.seed, ev.generation)
    _lib.check(L.ld_ga_step_device(C.byref(ga), ev.population.data_ptr(), rec.data_ptr(), n, ev.next_population.data_ptr(), 0, n, None))
    ev_[3].record()
    ev.population, ev.next_population = ev.next_population, ev.population
    ev.generation += 1
    torch.cuda.synchronize()
    acc += [ev_[0].elapsed_time(ev_[1]), ev_[1].elapsed_time(ev_[2]), ev_[2].elapsed_time(ev_[3]), 0]
wall = (time.perf_counter() - wall0) / G * 1e3
r = rec.cpu().numpy().view(np.int64)
ft = rec.view(torch.int32)[:, 3].cpu().numpy()
print(f"n={n}: per generation wall {wall:.3f} ms | fly {acc[0]/G:.3f} ms, reduce+sync {acc[1]/G:.3f} ms, ga {acc[2]/G:.3f} ms")
print(f"flight_time: mean {ft.mean():.1f}, p99 {np.quantile(ft, 0.99):.0f}, max {ft.max()} -> work-bound {ft.sum()/1.77e10*1e3:.3f} ms, longest episode {ft.max()*1.046e-3:.3f} ms")
