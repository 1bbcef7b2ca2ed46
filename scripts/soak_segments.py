#!/usr/bin/env python
"""Soak test of the segment scheduler (dev tool): random population sizes, segment lengths, refill batches and
horizons; every run must reproduce the unsegmented records bit for bit and finish (the kernel's own 20 s
guard turns a scheduling bug into an error instead of a hang)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from learning_drones_b200 import _lib, population as pop
from learning_drones_b200._abi import make_params
from oracle import gen_golden as G

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
wave = int(_lib.load().ld_wave_size())
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
pool_s = G.survivors(4 * wave, rng)
pool_r = G.random_brains(4 * wave, rng)
t0 = time.time()
for i in range(runs):
    n = int(rng.integers(wave + 1, 4 * wave))
    frac = rng.random()
    k = int(n * frac)
    genomes = np.concatenate([pool_s[:k], pool_r[: n - k]])[rng.permutation(n)]
    H = int(rng.integers(20, 400))
    p = make_params(H)
    gt = torch.from_numpy(genomes).cuda()
    p.slice_steps = -1
    ref, _ = pop.fly_population_device(gt, p)
    p.slice_steps = int(rng.integers(1, H + 20)) * (1 if rng.random() < 0.6 else -1)
    if p.slice_steps == -1:
        p.slice_steps = -2
    p.refill_batch = int(rng.integers(0, 33))
    if i % 3 == 0:
        out, _ = pop.fly_population_host(genomes, p)
        same = out.tobytes() == pop.records_to_numpy(ref).tobytes()
    else:
        out, _ = pop.fly_population_device(gt, p)
        same = bool(torch.equal(out, ref))
    if not same:
        print(f"MISMATCH run {i}: n={n} H={H} slice={p.slice_steps} batch={p.refill_batch}")
        sys.exit(1)
print(f"{runs} runs ok in {time.time() - t0:.1f} s")
