#!/usr/bin/env python
"""Per-step wall times of the blocking and the pipelined host API (dev tool, one GPU)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from learning_drones_b200 import population as pop
from learning_drones_b200._abi import RECORD_DTYPE, make_params

D, H = 65536, 2000
p = make_params(H)
p.slice_steps = int(os.environ.get("LD_SLICE_STEPS", "0"))
g = torch.from_numpy(bench.population("survivors", D)).pin_memory().numpy()
outs = [torch.empty((D, 4), dtype=torch.int64).pin_memory().numpy().view(RECORD_DTYPE).reshape(-1) for _ in range(2)]
for _ in range(3):
    pop.fly_population_host(g, p, out=outs[0])
pop.fly_population_host_submit(g, p, out=outs[1]).result()
t = [time.perf_counter()]
for k in range(8):
    pop.fly_population_host(g, p, out=outs[0]); t.append(time.perf_counter())
print("blocking  ms/step:", [round(1e3 * (b - a), 2) for a, b in zip(t, t[1:])])
t = [time.perf_counter()]
pending = None
for k in range(10):
    nxt = pop.fly_population_host_submit(g, p, out=outs[k & 1])
    ts = time.perf_counter()
    if pending is not None:
        pending.result()
    pending = nxt
    t.append(time.perf_counter())
pending.result(); t.append(time.perf_counter())
print("pipelined ms/step:", [round(1e3 * (b - a), 2) for a, b in zip(t, t[1:])], "total", round(1e3 * (t[-1] - t[0]), 2))
