#!/usr/bin/env python
"""Throughput of the secondary kernels (dev tool): PID-mode episodes and the GA step, with CUDA events."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from learning_drones_b200 import _lib, population as pop
from learning_drones_b200._abi import GENOME_LEN, make_ga_params, make_params, make_pid_params, paramset_to_pid_gains

L = _lib.load()
out = {}
# PID mode: 65,536 gain sets x 3,500-step collect_coins episodes
base = np.array([0.9, 0.9, 0.8, 0.1, 0.1, 0.1, 0.8, 0.8, 0.6, 0.9, 0.9, 0.2, 0.01, 0.01, 0.01, 0.12, 0.12, 0.12])
rng = np.random.default_rng(0)
pg = torch.from_numpy(paramset_to_pid_gains(base * rng.uniform(0.8, 1.2, size=(65536, 18)))).cuda()
p, q = make_params(), make_pid_params(3500)
for _ in range(2):
    rec, _ = pop.fly_pid_population_device(pg, p, q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    rec, _ = pop.fly_pid_population_device(pg, p, q)
e1.record()
torch.cuda.synchronize()
steps = int(pop.records_to_numpy(rec)["flight_time"].sum())
ms = e0.elapsed_time(e1) / 3
out["pid_mode"] = {"gain_sets": 65536, "ms": ms, "active_drone_steps": steps, "drone_steps_per_s": steps / (ms * 1e-3)}
# GA step: 1,048,576 drones (HBM-bound: 2 parent rows in, 2 child rows out per pair)
n = 1048576
parents = torch.randn((n, GENOME_LEN), dtype=torch.float64, device="cuda")
records = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
records[:, 0] = torch.randint(-5000, 4000, (n,), device="cuda")
children = torch.empty_like(parents)
ga = make_ga_params(1, 0)
for _ in range(2):
    _lib.check(L.ld_ga_step_device(C.byref(ga), parents.data_ptr(), records.data_ptr(), n, children.data_ptr(), 0, n, None))
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    _lib.check(L.ld_ga_step_device(C.byref(ga), parents.data_ptr(), records.data_ptr(), n, children.data_ptr(), 0, n, None))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
out["ga_step"] = {"drones": n, "ms": ms, "algorithmic_GBps": 2 * n * GENOME_LEN * 8 / (ms * 1e-3) / 1e9, "hbm_peak_GBps": 6533.5}
print(json.dumps(out))
