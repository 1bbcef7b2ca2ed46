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
