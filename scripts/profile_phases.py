#!/usr/bin/env python
"""Where does the episode kernel spend its cycles? (dev tool, one GPU)

Flies survivors / random / GA-offspring populations with the instrumented kernel (ld_debug_fly_profile_device) and
prints, per workload: kernel ms of the production kernel, share of warp-cycles per phase (claim / stage / wait / hot
loop / publish+park), cycles per staging event, lanes per event, live-lane utilisation of the hot loop.
    python scripts/profile_phases.py [drones] [timesteps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from learning_drones_b200 import population as pop
from learning_drones_b200._abi import make_params
from learning_drones_b200.evolve.gpu import GpuEvolution

D = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2000


def timed(gt, p, reps=5):
    rec = torch.empty((gt.shape[0], 4), dtype=torch.int64, device="cuda")
    cnt = torch.empty(2, dtype=torch.int64, device="cuda")
    for _ in range(2):
        pop.fly_population_device(gt, p, records=rec, work_counter=cnt)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        pop.fly_population_device(gt, p, records=rec, work_counter=cnt)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, int(pop.records_to_numpy(rec)["flight_time"].sum())


def report(name, gt, slice_steps, batch=0):
    p = make_params(H)
    p.slice_steps, p.refill_batch = slice_steps, batch
    ms, active = timed(gt, p)
    _, f = pop.fly_population_profile(gt, p)
    cyc = sum(f[k] for k in ("claim", "stage", "wait", "hot", "post"))
    share = {k: 100.0 * f[k] / cyc for k in ("claim", "stage", "wait", "hot", "post")}
    ev = max(f["stage_events"], 1)
    print(f"{name:34s} slice {slice_steps:5d} batch {batch:2d}: {ms:7.3f} ms  {active / ms / 1e6:7.2f} G active steps/s | "
          + " ".join(f"{k} {v:4.1f}%" for k, v in share.items())
          + f" | stage {f['stage'] / ev:7.0f} cyc/event (state {f['stage_state'] / ev:5.0f}, copy {f['stage_copy'] / ev:5.0f}, tmem {f['stage_tmem'] / ev:5.0f}, "
          + f"{f['stage_groups'] / ev:3.1f} groups), {f['lanes_staged'] / ev:4.1f} lanes/event, {ev / max(f['warps'], 1):6.1f} events/warp"
          + f" | hot: {f['hot'] / max(f['warp_steps'], 1):6.0f} cyc/warp-step, live lanes {100.0 * f['live_lane_steps'] / max(32 * f['warp_steps'], 1):4.1f}%"
          + f" | claim {f['claim'] / max(f['hot_entries'], 1):5.0f} post {f['post'] / max(f['hot_entries'], 1):5.0f} cyc/entry", flush=True)


surv = torch.from_numpy(bench.population("survivors", D)).cuda()
rnd = torch.from_numpy(bench.population("random", D)).cuda()
ev = GpuEvolution(bench.population("survivors", D), seed=2026, timesteps=H)
ev.run(2)
off2 = ev.population.clone()
ev.run(3)
off5 = ev.population.clone()
auto_ga = -max(125, H // 4)
QUICK = os.environ.get("LD_PROFILE_QUICK") == "1"
for name, gt, slices in (("survivors", surv, (0, -1)), ("random init", rnd, (0, -1)), ("GA offspring, generation 2", off2, (auto_ga, -1, -125)),
                         ("GA offspring, generation 5", off5, (auto_ga, -1, -125))):
    for sl in (slices[:2] if QUICK else slices):
        for batch in ((0,) if (name == "survivors" or QUICK) else (0, 2, 4, 8, 12)):
            report(name, gt, sl, batch)
