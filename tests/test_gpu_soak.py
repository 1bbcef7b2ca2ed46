"""Soak test of the episode kernel's work queue (park / re-queue / refill protocol), bounded to ~20 s.

compute-sanitizer is closed on this GPU pool (gpurun refuses it), so memory safety of the queue protocol is
exercised here instead: randomised population sizes, horizons, segment schedules and refill batches, device
and host entry points; every run must reproduce the unsegmented records bit for bit and finish (the kernel's
own wait limit turns a scheduling bug into an error instead of a hang)."""
import time

import numpy as np
import pytest

from learning_drones_b200._abi import make_params

pytestmark = pytest.mark.gpu
BUDGET_S = 20.0


def test_segment_queue_soak():
    import torch

    from learning_drones_b200 import _lib, population as pop
    from oracle import gen_golden as G

    wave = int(_lib.load().ld_wave_size())
    rng = np.random.default_rng(20261018)
    pool_s = G.survivors(4 * wave, rng)
    pool_r = G.random_brains(4 * wave, rng)
    t0, runs = time.time(), 0
    while time.time() - t0 < BUDGET_S or runs < 12:
        n = int(rng.integers(wave + 1, 4 * wave))
        k = int(n * rng.random())
        genomes = np.concatenate([pool_s[:k], pool_r[: n - k]])[rng.permutation(n)]
        H = int(rng.integers(20, 400))
        p = make_params(H)
        gt = torch.from_numpy(genomes).cuda()
        p.slice_steps = -1
        ref, _ = pop.fly_population_device(gt, p)
        p.slice_steps = int(rng.integers(1, H + 20)) * (1 if rng.random() < 0.6 else -1)
        if p.slice_steps == -1:
            p.slice_steps = -2
        if runs % 5 == 4:
            p.slice_steps = -1000000  # LD_SLICE_STAGGER
        p.refill_batch = int(rng.integers(0, 33))
        if runs % 3 == 0:
            out, _ = pop.fly_population_host(genomes, p)
            same = out.tobytes() == pop.records_to_numpy(ref).tobytes()
        else:
            out, _ = pop.fly_population_device(gt, p)
            same = bool(torch.equal(out, ref))
        assert same, f"run {runs}: n={n} H={H} slice={p.slice_steps} batch={p.refill_batch}"
        runs += 1
    print(f"soak: {runs} randomised runs bit-identical in {time.time() - t0:.1f} s")


def test_soak_with_the_kernels_own_bounds_checks():
    """The same kind of run through libldrones_checked.so (-DLD_CHECKED): slot / drone / re-queue / publish / staging
    indices are range-checked inside the kernel and a violation is reported through the launch's error word.  Runs in a
    subprocess because a process binds one build of the library."""
    import os
    import subprocess
    import sys

    from learning_drones_b200 import build as ldbuild

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, ROOT_DIR)
from learning_drones_b200 import _lib, population as pop
from learning_drones_b200._abi import make_params
from oracle import gen_golden as G
assert _lib.LIB_PATH.endswith("libldrones_checked.so")
wave = int(_lib.load().ld_wave_size())
rng = np.random.default_rng(77)
pool = np.concatenate([G.survivors(2 * wave, rng), G.random_brains(2 * wave, rng)])
for run in range(10):
    n = int(rng.integers(wave + 1, 4 * wave))
    g = pool[rng.permutation(len(pool))[:n]]
    H = int(rng.integers(30, 300))
    p = make_params(H)
    p.slice_steps = -1
    ref, _ = pop.fly_population_host(g, p)
    p.slice_steps = -1000000 if run % 4 == 3 else int(rng.integers(2, H)) * (1 if run % 2 else -1)   # -1000000 = LD_SLICE_STAGGER
    p.refill_batch = int(rng.integers(0, 33))
    out, _ = pop.fly_population_host(g, p)          # raises LdError if the kernel's error word is set
    assert out.tobytes() == ref.tobytes(), run
    gt = torch.from_numpy(g).cuda()
    cnt = torch.empty(2, dtype=torch.int64, device="cuda")
    rec, _ = pop.fly_population_device(gt, p, work_counter=cnt)
    torch.cuda.synchronize()
    pop.check_work_counter(cnt)
    assert pop.records_to_numpy(rec).tobytes() == ref.tobytes(), run
print("checked soak ok")
""".replace("ROOT_DIR", repr(root))
    env = dict(os.environ, LDRONES_LIB=ldbuild.LIB_CHECKED)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "checked soak ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
