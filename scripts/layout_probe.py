#!/usr/bin/env python
"""Cycles per warp-step of the episode kernel's CTA layouts (dev tool, one GPU): flies exactly one wave of survivors,
unsegmented, so that the kernel time is the slowest warp's 2000 steps, and reads the instrumented build's average.
    LD_WARPS_PER_SM=5 python scripts/layout_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from learning_drones_b200 import _lib
from learning_drones_b200 import population as pop
from learning_drones_b200._abi import make_params

H = 2000
wave = int(_lib.load().ld_wave_size())
for n in (wave, 2 * wave, 65536):
    gt = torch.from_numpy(bench.population("survivors", n)).cuda()
    p = make_params(H)
    p.slice_steps = -1
    rec = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    cnt = torch.empty(2, dtype=torch.int64, device="cuda")
    for _ in range(2):
        pop.fly_population_device(gt, p, records=rec, work_counter=cnt)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        pop.fly_population_device(gt, p, records=rec, work_counter=cnt)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    _, f = pop.fly_population_profile(gt, p)
    print(f"LD_WARPS_PER_SM={os.environ.get('LD_WARPS_PER_SM', '4')} wave {wave} n {n}: {ms:.3f} ms unsegmented = {ms * 1e-3 * 1.965e9 / H / ((n + wave - 1) // wave):.0f} cycles per step "
          f"of the slowest lane; instrumented average {f['hot'] / max(f['warp_steps'], 1):.0f} cycles per warp-step over {f['warps']} warps; "
          f"{n * H / ms / 1e6:.2f} G drone-steps/s", flush=True)
