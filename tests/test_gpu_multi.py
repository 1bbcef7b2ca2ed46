"""Multi-GPU checks that need >= 2 visible GPUs (skipped on a single-GPU box): the NVLink
peer-memory GA step against the replicated one, and bench.py under torchrun."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count()


def _torchrun(args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_p2p_evolution_matches_replicated():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(["scripts/p2p_evolution_check.py"], 29541)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "p2p == replicated: True" in r.stdout


def test_bench_two_ranks():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(["bench.py", "--gpus", "2", "--steps", "3", "--generations", "2"], 29542)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["value"] > 1e10
    assert line["config"]["global_drones"] == 2 * line["config"]["drones_per_gpu"]
