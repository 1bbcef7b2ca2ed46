#!/usr/bin/env python
"""A/B check of two builds of libldrones.so (dev tool): fly the same populations and dump the records + final states so
that two runs (LDRONES_LIB=... python scripts/ab_records.py out.npz) can be compared bit for bit with --compare."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    ok = True
    for k in a.files:
        same = a[k].tobytes() == b[k].tobytes()
        ok &= same
        print(k, "identical" if same else f"DIFFERENT ({int((a[k] != b[k]).sum())} entries)")
    sys.exit(0 if ok else 1)

import torch

import bench
from learning_drones_b200 import population as pop
from learning_drones_b200._abi import make_params

out = {}
for kind, n, H in (("random", 131072, 2000), ("survivors", 40000, 700)):
    g = torch.from_numpy(bench.population(kind, n)).cuda()
    rec, st = pop.fly_population_device(g, make_params(H), want_state=True)
    out[f"{kind}_rec"] = rec.cpu().numpy()
    out[f"{kind}_state"] = st.cpu().numpy().view(np.int64)
# saturating / non-finite brains exercise the exp clamp and the overflow test
rng = np.random.default_rng(1)
g = rng.normal(0, 1, size=(4096, 314)) * rng.choice([1.0, 30.0, 1e3, 1e6, 1e150, 1e300], size=(4096, 1))
rec, st = pop.fly_population_device(torch.from_numpy(g).cuda(), make_params(600), want_state=True)
out["wild_rec"], out["wild_state"] = rec.cpu().numpy(), st.cpu().numpy().view(np.int64)
np.savez(sys.argv[1], **out)
print("written", sys.argv[1])
