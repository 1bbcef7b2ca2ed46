#!/usr/bin/env python
"""generations/s of GpuEvolution on one GPU: generations enqueued in one block vs one host sync each (dev tool)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util as iu

import torch

sp = iu.spec_from_file_location("ldbench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
bench = iu.module_from_spec(sp)
sp.loader.exec_module(bench)
from learning_drones_b200.evolve.gpu import GpuEvolution

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
G = int(sys.argv[2]) if len(sys.argv) > 2 else 5
genomes = bench.population("survivors", n)
for mode in ("block", "step", "block", "step"):
    ev = GpuEvolution(genomes, seed=2026, timesteps=2000, **({"slice_steps": int(os.environ["LD_SLICE_STEPS"])} if "LD_SLICE_STEPS" in os.environ else {}))
    ev.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hist = ev.run(G) if mode == "block" else [ev.step() for _ in range(G)]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{mode}: {G} generations in {dt*1e3:.2f} ms = {G/dt:.1f} generations/s; steps {[h['active_steps'] for h in hist][:3]}")
