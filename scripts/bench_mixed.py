#!/usr/bin/env python
"""Kernel time of 65,536-drone populations with a given fraction of survivors (the rest: random brains that
mostly crash early), for several segment lengths (dev tool).  usage: bench_mixed.py [fractions...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from learning_drones_b200 import population as pop
from learning_drones_b200._abi import make_params
from oracle import gen_golden as G

fracs = [float(a) for a in sys.argv[1:]] or [0.25, 0.5, 0.9, 1.0]
rng = np.random.default_rng(3)
n = 65536
for frac in fracs:
    k = int(n * frac)
    g = np.concatenate([G.survivors(k, rng), G.random_brains(n - k, rng)])[rng.permutation(n)]
    gt = torch.from_numpy(g).cuda()
    row = []
    for sl in (0, 500, 250, -250, -125, -500):
        p = make_params(2000)
        p.slice_steps = sl
        rec = None
        for _ in range(2):
            rec, _ = pop.fly_population_device(gt, p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            pop.fly_population_device(gt, p, records=rec)
        e1.record()
        torch.cuda.synchronize()
        row.append(f"slice {sl}: {e0.elapsed_time(e1) / 5:.3f}")
    print(f"survivors {frac:4.2f} | " + " | ".join(row) + " ms")
