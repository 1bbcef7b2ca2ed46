#!/usr/bin/env python
"""Kernel ms of the segment schedules on the populations of the bench's GA leg, generation by generation (dev tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from learning_drones_b200 import population as pop
from learning_drones_b200._abi import SLICE_STAGGER, make_params
from learning_drones_b200.evolve.gpu import GpuEvolution

D, H = 65536, 2000
ev = GpuEvolution(bench.population("survivors", D), seed=2026, timesteps=H)


def timed(gt, sl, reps=4):
    p = make_params(H)
    p.slice_steps = sl
    rec = torch.empty((D, 4), dtype=torch.int64, device="cuda")
    cnt = torch.empty(2, dtype=torch.int64, device="cuda")
    pop.fly_population_device(gt, p, records=rec, work_counter=cnt)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        pop.fly_population_device(gt, p, records=rec, work_counter=cnt)
    b.record()
    torch.cuda.synchronize()
    r = pop.records_to_numpy(rec)
    return a.elapsed_time(b) / reps, int((r["flight_time"] >= H).sum()), int(r["flight_time"].sum())


for g in range(7):
    gt = ev.population.clone()
    row = []
    for name, sl in (("off", -1), ("dbl500", -500), ("dbl250", -250), ("stagger", SLICE_STAGGER)):
        ms, full, steps = timed(gt, sl)
        row.append(f"{name} {ms:6.3f}")
    print(f"generation {g}: {full:6d} full-horizon drones, {steps / 1e6:6.1f} M steps, bound {max(steps / 18944 * 1.0e-3 * 0.99, 2.0 if full else 0):5.2f} ms | " + " | ".join(row), flush=True)
    ev.step()
