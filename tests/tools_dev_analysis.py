"""GPU-vs-oracle deviation as a function of the oracle's conditioning number (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from _parity import conditioning, state_dev
from learning_drones_b200 import population as pop
from learning_drones_b200._abi import make_params
from oracle import oracle as O, gen_golden as G
rng = np.random.default_rng(2026)
G.survivors(65536, rng)
g = G.random_brains(65536, rng)[:20000]
p = make_params(2500)
cond = conditioning(O, p, g)
ref, rst = O.fly_mlp(p, g, want_state=True)
rec, st = pop.fly_population_host(g, p, want_state=True)
dev = state_dev(st, rst)
same = (rec["reward"] == ref["reward"]) & (rec["flight_time"] == ref["flight_time"]) & (rec["coins"] == ref["coins"])
edges = [0, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-6, 1e-3, np.inf]
for lo, hi in zip(edges[:-1], edges[1:]):
    m = (cond >= lo) & (cond < hi)
    if m.any():
        print(f"cond [{lo:.0e},{hi:.0e}): n={m.sum():6d} ints_same={same[m].mean():.4f} dev max={dev[m].max():.2e} med={np.median(dev[m]):.2e} ratio max={np.max(dev[m]/np.maximum(cond[m],1e-17)):.1e}")
print("overall ints same", same.mean(), "n diff", (~same).sum())
