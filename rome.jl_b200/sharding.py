"""Factor-list sharding across ranks: contiguous equal-size ranges of the factor list, particles
replicated, one all-gather of the proposal rows per sweep.  Pure host-side plumbing over torch.distributed
(NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_size(n_factors: int, world: int) -> int:
    """rows per rank; the row buffers are allocated for world * shard_size(...) factors"""
    return -(-n_factors // world)


def shard_range(n_factors: int, rank: int, world: int):
    """(first, count) of the factors rank `rank` evaluates; count may be 0 for trailing ranks"""
    c = shard_size(n_factors, world)
    first = min(rank * c, n_factors)
    return first, max(0, min(c, n_factors - first))


def allgather_rows(rows, n_factors: int, group=None):
    """In-place all-gather of a globally indexed row buffer [world*shard][...]: every rank contributes the slice it
    evaluated (its shard_range) and receives everybody else's."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    c = shard_size(n_factors, world)
    if rows.shape[0] != world * c:
        raise ValueError(f"row buffer must hold world*shard = {world * c} factors, has {rows.shape[0]}")
    mine = rows[rank * c:(rank + 1) * c]
    try:
        dist.all_gather_into_tensor(rows, mine, group=group)
    except (RuntimeError, NotImplementedError):  # backend without the flat variant
        parts = [rows[r * c:(r + 1) * c] for r in range(world)]
        dist.all_gather(parts, mine.clone(), group=group)
    return rows
