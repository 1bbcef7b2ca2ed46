#!/usr/bin/env python
"""Can a child's episode length be predicted from its parents'? (dev tool, one GPU)

For generations of the bench's GA leg (65,536 survivors start population): flies generation g, derives every
child's tournament winner with the oracle's GA restatement (same Philox draws as the kernel), flies generation g+1
and reports how well "winner flew the whole horizon" predicts "child flies the whole horizon" — the information a
longest-first queue order would need.  Also simulates list scheduling (lanes = one wave) in index order vs
predicted-long-first order vs clairvoyant longest-first order.
"""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from learning_drones_b200 import _lib
from learning_drones_b200._abi import make_ga_params
from learning_drones_b200.evolve.gpu import GpuEvolution
from learning_drones_b200.population import fly_population_device, records_to_numpy
from oracle import oracle as O

D, H = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 2000
LANES = int(_lib.load().ld_wave_size())


def makespan(order, life):
    lanes = [0] * min(LANES, len(order))
    heapq.heapify(lanes)
    end = 0
    for i in order:
        t = heapq.heappop(lanes) + int(life[i])
        end = max(end, t)
        heapq.heappush(lanes, t)
    return end


ev = GpuEvolution(bench.population("survivors", D), seed=2026, timesteps=H)
prev = None
for g in range(7):
    pop_g = ev.population.clone()
    rec = records_to_numpy(fly_population_device(pop_g, ev.params)[0])
    life = rec["flight_time"].astype(np.int64)
    if prev is not None:
        winner_life, other_life = prev
        full = life >= H
        pred = winner_life >= H
        both = pred & (other_life >= H)
        tp = int((pred & full).sum())
        print(f"gen {g}: {int(full.sum())} of {D} fly the whole horizon ({life.sum() / 1e6:.1f} M steps); predictor 'winner flew it all': "
              f"predicted {int(pred.sum())}, precision {tp / max(pred.sum(), 1):.2f}, recall {tp / max(full.sum(), 1):.2f}; "
              f"'both pair parents': predicted {int(both.sum())}, precision {int((both & full).sum()) / max(both.sum(), 1):.2f}, "
              f"recall {int((both & full).sum()) / max(full.sum(), 1):.2f}; spearman-ish corr(winner life, life) = "
              f"{np.corrcoef(winner_life, life)[0, 1]:.2f}")
        idx = np.arange(D)
        by_pred = np.concatenate([idx[pred], idx[~pred]])
        by_parent = np.argsort(-winner_life, kind="stable")
        clair = np.argsort(-life, kind="stable")
        bound = max(int(life.max()), int(np.ceil(life.sum() / LANES)))
        print(f"        list-scheduling makespan in steps ({LANES} lanes): index order {makespan(idx, life)}, predicted-long first "
              f"{makespan(by_pred, life)}, sorted by winner's life {makespan(by_parent, life)}, clairvoyant longest-first {makespan(clair, life)}, bound {bound}", flush=True)
    # children of this generation (the kernel's draws, restated on the CPU)
    ga = make_ga_params(ev.seed, ev.generation, **ev.hyper)
    _, picks = O.ga_step(ga, pop_g.cpu().numpy(), rec["reward"].astype(np.int64))
    winner = picks[:, 0]
    partner = picks[np.arange(D) ^ 1, 0]
    prev = (life[winner].copy(), life[partner].copy())
    prev[0][0] = life[int(np.argmax(rec["reward"]))]  # slot 0 is the elite (approximately: the generation's best)
    ev.step()
