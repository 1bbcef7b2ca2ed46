#!/usr/bin/env python
"""When does every drone of a GA-offspring generation get its lane, and when does it land? (dev tool, one GPU)

Flies the bench's generation-5 population with the instrumented kernel's timeline tap and compares the measured
starts / landings with an ideal greedy list schedule of the same lifetimes on the same number of lanes."""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from learning_drones_b200 import _lib, population as pop
from learning_drones_b200._abi import make_params
from learning_drones_b200.evolve.gpu import GpuEvolution

D, H = 65536, 2000
LANES = int(_lib.load().ld_wave_size())
ev = GpuEvolution(bench.population("survivors", D), seed=2026, timesteps=H)
ev.run(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
gt = ev.population.clone()
for slice_steps in (-1, -500, -125):
    p = make_params(H)
    p.slice_steps = slice_steps
    rec, f = pop.fly_population_profile(gt, p, timeline=True)
    life = pop.records_to_numpy(rec)["flight_time"].astype(np.int64)
    t = f["times"].astype(np.float64) / 1e3  # us
    start, land = t[:, 0], t[:, 1]
    order = np.argsort(start, kind="stable")
    step_us = np.median((land - start)[life >= H] / H) if (life >= H).any() else 1.03
    # ideal greedy schedule in the measured start order
    lanes = [0.0] * LANES
    heapq.heapify(lanes)
    ideal_start = np.zeros(D)
    for i in order:
        s = heapq.heappop(lanes)
        ideal_start[i] = s
        heapq.heappush(lanes, s + life[i] * step_us)
    ideal_end = (ideal_start + life * step_us).max()
    full = life >= H
    q = lambda a: np.percentile(a, [50, 90, 99, 100]).round(0).tolist()
    print(f"slice {slice_steps}: kernel span {land.max():.0f} us; step {step_us:.3f} us; {int(full.sum())} full-horizon drones; "
          f"last start measured {start.max():.0f} us vs ideal greedy {ideal_start.max():.0f} us; ideal makespan {ideal_end:.0f} us")
    print(f"   start delay vs ideal (us) p50/p90/p99/max: {q(start - ideal_start)}; full-horizon drones: start p50/p90/max {q(start[full])}, "
          f"flight duration / (H x step) p50/p90/max {q((land - start)[full] / (H * step_us) * 100)} %")
    # lane occupancy over time: drones in flight (between first start and landing) per 100 us
    edges = np.arange(0, land.max() + 100, 100)
    occ = [(int(((start <= e) & (land > e)).sum())) for e in edges]
    print("   drones between start and landing every 100 us:", occ, flush=True)
    print(f"   warp-steps {f['warp_steps']}, of them through the large-angle libm trigonometry: {f['slow_trig_steps']} ({100.0 * f['slow_trig_steps'] / max(f['warp_steps'], 1):.2f} %)")
    if slice_steps == -1:
        _, st = pop.fly_population_device(gt, p, want_state=True)
        ang = np.abs(st.cpu().numpy()[:, 6:9]).max(axis=1)
        status = pop.records_to_numpy(rec)["status"]
        print(f"   non-finite drones (status bit 2): {int(((status & 2) != 0).sum())}, of them flying the whole horizon: {int((((status & 2) != 0) & full).sum())}; "
              f"non-finite final angles: {int((~np.isfinite(ang)).sum())}")
        big = ang > 105615.0
        dur = land - start
        print(f"   drones that end beyond the fast trigonometric range (|angle| > 1.05e5 rad): {int(big.sum())} of {D}, {int((big & full).sum())} of them "
              f"full-horizon; |angle| quantiles of full-horizon drones p50/p90/p99/max {q(ang[full])}; stretch of full-horizon drones with / without a "
              f"big final angle: {np.median(dur[full & big]) / (H * step_us) if (full & big).any() else float('nan'):.3f} / {np.median(dur[full & ~big]) / (H * step_us):.3f}", flush=True)
