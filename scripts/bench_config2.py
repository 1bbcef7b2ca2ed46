#!/usr/bin/env python
"""BASELINE config 2: the reference's default evolution (20 drones, seed 12345, H = 2,500) through the drop-in
API — `Drone()`, `fly_population(drones)`, `evolve.util.selection_and_crossover`, `Drone.mutate`, elitism —
i.e. evolve/evolution.py:40-118 with the per-drone fly() loop replaced by one library call per generation.
Prints generations/s and the time of the fitness evaluation alone (dev tool, one GPU)."""
import os
import sys
import time
from copy import deepcopy

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_drones_b200.drone.drone import Drone
from learning_drones_b200.evolve.util import selection_and_crossover
from learning_drones_b200.flight import fly_population

n_generations = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_drones, mutation_rate, p_tournament, p_crossover = 20, 0.05, 0.6, 0.05  # evolution.py:16-20

np.random.seed(12345)
drones = [Drone() for _ in range(n_drones)]
fly_population(deepcopy(drones))  # warm-up: library load, workspace
best_score, best = -np.inf, None
t_fly = 0.0
t0 = time.perf_counter()
steps = 0
for gen in range(n_generations):
    a = time.perf_counter()
    results = fly_population(drones)
    t_fly += time.perf_counter() - a
    scores = np.array([r[0] for r in results])
    steps += sum(d.flight_time for d in drones)
    i = int(np.argmax(scores))
    if scores[i] > best_score:
        best_score, best = scores[i], deepcopy(drones[i])
    new_population = selection_and_crossover(drones, p_tournament, p_crossover)
    for d in new_population:
        d.reset()
        d.mutate(mutation_rate)
    drones = new_population
    drones[0] = best
    drones[0].reset()
dt = time.perf_counter() - t0
print(f"config 2: {n_generations} generations of {n_drones} drones in {dt:.3f} s = {n_generations/dt:.1f} generations/s "
      f"(fitness evaluation {1e3*t_fly/n_generations:.2f} ms per generation, {steps} active drone-steps, best score {best_score})")
