"""World-size-2 test of the multi-GPU host logic on CPU (gloo): shard -> evaluate -> all-gather.

The per-shard evaluation is done by the CPU oracle here (no GPU in this container); the code
under test is the sharding arithmetic and the record all-gather that bench.py / evolution use
with NCCL on the GPU box."""
import os
import socket

import numpy as np
import pytest

from learning_drones_b200._abi import RECORD_DTYPE, make_params
from learning_drones_b200.distributed import shard_bounds, shard_sizes


def test_shard_bounds_cover_population():
    for n in (0, 1, 7, 20, 65536, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gen_golden as G
    from oracle import oracle as O
    from learning_drones_b200.distributed import allgather_records, records_view

    genomes = G.random_brains(n, np.random.default_rng(99))
    lo, hi = shard_bounds(n, rank, world)
    p = make_params(400)
    rec, _ = O.fly_mlp(p, genomes[lo:hi], nthreads=1)
    local = torch.from_numpy(rec.view(np.int64).reshape(-1, 4).copy())
    gathered = allgather_records(local, n_global=n)
    gathered2 = allgather_records(local)  # sizes discovered with a collective
    assert torch.equal(gathered, gathered2)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), records_view(gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 61])  # equal and ragged shards
def test_allgather_records_world2(tmp_path, oracle, n):
    import torch.multiprocessing as mp

    from oracle import gen_golden as G

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    genomes = G.random_brains(n, np.random.default_rng(99))
    full, _ = oracle.fly_mlp(make_params(400), genomes, nthreads=1)
    for r in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{r}.npy"))
        assert got.dtype == RECORD_DTYPE and got.tobytes() == full.tobytes()
