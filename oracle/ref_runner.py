"""Time the REFERENCE's own fly() (unmodified files staged under oracle/_ref by oracle/build_ref.sh, or the tree at
$LD_REFERENCE / /root/reference when present) on the host cores — TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

BASELINE.md §3: `multiprocessing.Pool(nproc)`, one drone per task, weights injected into fresh `Drone` objects,
`visualize=False`, `time.perf_counter()` around the whole map, active drone-steps = sum of flight_time.
Shims: `np.Infinity = np.inf` (numpy >= 2), scratch cwd (every Drone() creates
./evolution/<timestamp>/), stdout silenced (Drone() prints an [INFO] line).
"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LAYER_SHAPES = ((10, 16), (10, 11), (4, 11))


def reference_root():
    """Directory holding the reference's packages, or None."""
    for cand in (os.path.join(HERE, "_ref"), os.environ.get("LD_REFERENCE", ""), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "flight", "__init__.py")) and os.path.isfile(os.path.join(cand, "drone", "drone.py")):
            return cand
    return None


_state = {}


def _init_worker(root, scratch):
    np.Infinity = np.inf
    if root not in sys.path:
        sys.path.insert(0, root)
    os.chdir(scratch)
    import logging

    logging.disable(logging.CRITICAL)
    with contextlib.redirect_stdout(io.StringIO()):
        from drone.drone import Drone
        from flight import fly
    _state["Drone"], _state["fly"] = Drone, fly


def _fly_one(args):
    genome, timesteps = args
    Drone, fly = _state["Drone"], _state["fly"]
    with contextlib.redirect_stdout(io.StringIO()):
        d = Drone(visualize=False, n_servers=0)
    ws, o = [], 0
    for r, c in LAYER_SHAPES:
        ws.append(np.array(genome[o:o + r * c], dtype=np.float64).reshape(r, c))
        o += r * c
    d.controller.weights = ws
    try:
        reward, _, _ = fly(d, timesteps=timesteps, idx=0)
    except ValueError:  # int(nan), drone/drone.py:14
        reward = d.reward
    return int(reward), int(d.flight_time), int(d.coins), bool(d.crashed)


class ReferencePool:
    """A pool of worker processes that have the reference imported; `fly(genomes, timesteps)` -> dict."""

    def __init__(self, nproc=None):
        import multiprocessing as mp

        self.root = reference_root()
        if self.root is None:
            raise FileNotFoundError("reference files not found: run oracle/build_ref.sh where /root/reference exists")
        self.nproc = nproc or len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (nproc or os.cpu_count())
        self.scratch = tempfile.mkdtemp(prefix="ld_ref_")
        self.pool = mp.get_context("fork").Pool(self.nproc, initializer=_init_worker, initargs=(self.root, self.scratch))

    def fly(self, genomes, timesteps):
        g = np.ascontiguousarray(genomes, dtype=np.float64)
        t0 = time.perf_counter()
        res = self.pool.map(_fly_one, [(row, int(timesteps)) for row in g], chunksize=1)
        dt = time.perf_counter() - t0
        reward, flight_time, coins, crashed = (np.array(x) for x in zip(*res))
        return dict(reward=reward, flight_time=flight_time, coins=coins, crashed=crashed, seconds=dt,
                    active_steps=int(flight_time.sum()), cores=self.nproc)

    def close(self):
        self.pool.terminate()
        self.pool.join()
        import shutil

        shutil.rmtree(self.scratch, ignore_errors=True)
