#!/usr/bin/env python
"""Phase profile of the episode kernel on a config-5-like population (dev tool, one GPU): 131,072 reference-init drones after
20 generations of the GPU loop — every drone crashes, mean life ~170 steps — flown unsegmented with the instrumented build."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from learning_drones_b200 import population as pop
from learning_drones_b200._abi import make_params
from learning_drones_b200.evolve.gpu import GpuEvolution

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
H = 2000
rng = np.random.default_rng(0)
genomes = np.concatenate([rng.normal(0.0, 1.0 / np.sqrt(r), size=(n, r * c)) for r, c in ((10, 16), (10, 11), (4, 11))], axis=1)
ev = GpuEvolution(genomes, seed=0, timesteps=H)
ev.run(20)
gt = ev.population.clone()
for sl, batch in ((-1, 0), (-1, 4), (-1, 8), (-500, 0)):
    p = make_params(H)
    p.slice_steps, p.refill_batch = sl, batch
    rec = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    cnt = torch.empty(2, dtype=torch.int64, device="cuda")
    for _ in range(2):
        pop.fly_population_device(gt, p, records=rec, work_counter=cnt)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        pop.fly_population_device(gt, p, records=rec, work_counter=cnt)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    life = pop.records_to_numpy(rec)["flight_time"].astype(np.int64)
    _, f = pop.fly_population_profile(gt, p)
    cyc = sum(f[k] for k in ("claim", "stage", "wait", "hot", "post"))
    ev_ = max(f["stage_events"], 1)
    print(f"slice {sl} batch {batch}: {ms:.3f} ms, {life.sum() / ms / 1e6:.2f} G active steps/s (life mean {life.mean():.0f} max {life.max()}) | "
          + " ".join(f"{k} {100.0 * f[k] / cyc:4.1f}%" for k in ("claim", "stage", "wait", "hot", "post"))
          + f" | stage {f['stage'] / ev_:.0f} cyc/event, {f['lanes_staged'] / ev_:.1f} lanes/event, {ev_ / max(f['warps'], 1):.1f} events/warp"
          + f" | hot {f['hot'] / max(f['warp_steps'], 1):.0f} cyc/warp-step, live lanes {100.0 * f['live_lane_steps'] / max(32 * f['warp_steps'], 1):.1f}%, "
          + f"{f['warp_steps'] / max(f['hot_entries'], 1):.1f} steps/entry | claim {f['claim'] / max(f['hot_entries'], 1):.0f} post {f['post'] / max(f['hot_entries'], 1):.0f} cyc/entry",
          flush=True)
