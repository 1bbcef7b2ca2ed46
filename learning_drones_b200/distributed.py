"""Multi-GPU plumbing: one process per GPU, the population sharded in contiguous blocks.

Individuals are independent inside fly() (the reference evaluates them one by one,
evolve/evolution.py:64), so the data path needs no collective.  The one exchange step per
generation is an all-gather of the 32-byte fitness records so that every rank can run
selection on the whole population (NCCL over NVLink on the GPU box; the same code runs on
gloo/CPU tensors in the tests).  torch.distributed is plumbing only.
"""
import numpy as np

from ._abi import RECORD_DTYPE

RECORD_WORDS = RECORD_DTYPE.itemsize // 8


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block [lo, hi) of an n-drone population owned by `rank`; sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int):
    return [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]


def allgather_records(local_records, n_global: int = None, group=None):
    """All-gather per-rank record tensors (int64 [n_local, 4]) into population order.

    Equal shards use one all_gather_into_tensor (a single NCCL all-gather); ragged shards are
    padded to the largest shard and trimmed.  Returns an int64 [n_global, 4] tensor on the same
    device as `local_records`.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local_records
    world = dist.get_world_size(group)
    n_local = local_records.shape[0]
    if n_global is None:
        sizes_t = torch.tensor([n_local], dtype=torch.int64, device=local_records.device)
        all_sizes = [torch.zeros_like(sizes_t) for _ in range(world)]
        dist.all_gather(all_sizes, sizes_t, group=group)
        sizes = [int(s.item()) for s in all_sizes]
    else:
        sizes = shard_sizes(n_global, world)
        if sizes[dist.get_rank(group)] != n_local:
            raise ValueError(f"local shard has {n_local} records, expected {sizes[dist.get_rank(group)]}")
    local_records = local_records.contiguous()
    if len(set(sizes)) == 1:
        out = torch.empty((world * n_local, RECORD_WORDS), dtype=local_records.dtype, device=local_records.device)
        dist.all_gather_into_tensor(out, local_records, group=group)
        return out
    m = max(sizes)
    padded = torch.zeros((m, RECORD_WORDS), dtype=local_records.dtype, device=local_records.device)
    padded[:n_local] = local_records
    out = torch.empty((world * m, RECORD_WORDS), dtype=local_records.dtype, device=local_records.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * m : r * m + sizes[r]] for r in range(world)], dim=0)


def evaluate_generation(local_genomes, params, n_global: int = None, group=None):
    """Fly this rank's shard on its GPU and return (local_records, all_records): the fitness
    evaluation of one generation of evolve/evolution.py, population-parallel."""
    from .population import fly_population_device

    local_records, _ = fly_population_device(local_genomes, params)
    return local_records, allgather_records(local_records, n_global, group)


def records_view(records) -> np.ndarray:
    """int64 [n,4] tensor (any device) -> structured numpy records."""
    return records.detach().cpu().numpy().view(RECORD_DTYPE).reshape(-1)
