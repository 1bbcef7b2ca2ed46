"""GA operators with the reference's semantics and RNG draw order (evolve/util.py:5-104).

These are the host-side (object-level) operators used by the reference's evolution loop; they
run between fly_population() calls and are not part of the per-timestep hot path.
"""
from copy import deepcopy

import numpy as np


def tournament_select(drones, idx1, idx2, p_tournament):
    r = np.random.random_sample()
    better, worse = (idx1, idx2) if drones[idx1].reward > drones[idx2].reward else (idx2, idx1)
    return better if r < p_tournament else worse


def crossover(weights1, weights2, p_crossover):
    """Per layer, with probability p_crossover, single-point crossover of the C-order-flattened layer."""
    new1, new2 = [], []
    for i, layer in enumerate(weights1):
        if np.random.random_sample() < p_crossover:
            rows = layer.shape[0]
            cut = np.random.randint(0, rows * layer.shape[1])
            a, b = layer.flatten(), weights2[i].flatten()
            new1.append(np.concatenate([a[:cut], b[cut:]]).reshape(rows, -1))
            new2.append(np.concatenate([b[:cut], a[cut:]]).reshape(rows, -1))
        else:
            new1.append(weights1[i])
            new2.append(weights2[i])
    return new1, new2


def selection_and_crossover(drones, p_tournament, p_crossover):
    n = len(drones)
    new_population = deepcopy(drones)
    for i in range(0, n, 2):
        idx1, idx2, idx3, idx4 = (np.random.randint(0, n) for _ in range(4))
        s1 = tournament_select(drones, idx1, idx2, p_tournament)
        s2 = tournament_select(drones, idx3, idx4, p_tournament)
        new_population[i] = deepcopy(drones[s1])
        new_population[i + 1] = deepcopy(drones[s2])
        w1, w2 = crossover(new_population[i].controller.weights, new_population[i + 1].controller.weights, p_crossover)
        new_population[i].controller.weights = w1
        new_population[i + 1].controller.weights = w2
    return new_population


def get_best_of_generation(scores):
    best_index, best_score = -1, -np.inf
    for i, score in enumerate(scores):
        if score > best_score:
            best_index, best_score = i, score
    return best_index, best_score


def compare_weights(weights1, weights2):
    from ..drone.drone import Drone

    if isinstance(weights1, Drone):
        weights1 = weights1.controller.weights
    if isinstance(weights2, Drone):
        weights2 = weights2.controller.weights
    for i, layer in enumerate(weights1):
        for j, row in enumerate(layer):
            for k, w in enumerate(row):
                assert w == weights2[i][j][k], f"i: {i}, j: {j}, k: {k}, w1: {w}, w2: {weights2[i][j][k]}"
