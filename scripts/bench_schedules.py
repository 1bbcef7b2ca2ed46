#!/usr/bin/env python
"""Kernel ms of segment schedules on survivor-fraction mixes (dev tool, one GPU): unsegmented / automatic (doubling from
H/4) / LD_SLICE_STAGGER, for several population sizes; plus the end-to-end ms of the blocking host call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from learning_drones_b200 import population as pop
from learning_drones_b200._abi import RECORD_DTYPE, SLICE_STAGGER, make_params

H = 2000


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
    return a.elapsed_time(b) / reps


def e2e(g, p, reps=5):
    out = torch.empty((g.shape[0], 4), dtype=torch.int64).pin_memory().numpy().view(RECORD_DTYPE).reshape(-1)
    for _ in range(2):
        pop.fly_population_host(g, p, out=out)
    t0 = time.perf_counter()
    for _ in range(reps):
        pop.fly_population_host(g, p, out=out)
    return 1e3 * (time.perf_counter() - t0) / reps


for D in (65536, 100000):
    s, r = bench.population("survivors", D), bench.population("random", D)
    rng = np.random.default_rng(3)
    for frac in (1.0, 0.9):
        g = s.copy()
        mask = rng.random(D) >= frac
        g[mask] = r[mask]
        pinned = torch.from_numpy(g).pin_memory()
        gt = pinned.cuda()
        row = []
        for name, sl in (("off", -1), ("auto", 0), ("stagger", SLICE_STAGGER)):
            p = make_params(H)
            p.slice_steps = sl
            row.append(f"{name} {timed(gt, p):7.3f} ms (e2e {e2e(pinned.numpy(), p):7.3f})")
        print(f"{D:7d} drones, {100 * frac:3.0f} % survivors: " + " | ".join(row), flush=True)
