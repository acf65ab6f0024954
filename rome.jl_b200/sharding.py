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


# ---- owner partition (the exchange that replaces the all-gather; host planning only so far) ------------
def owner_of(index, nvars: int, world: int):
    """owner rank of variable `index` when the `nvars` variables of a type are split into `world` contiguous ranges of
    shard_size(nvars, world) (the last ranks may own fewer, or none)"""
    import numpy as np
    return np.minimum(np.asarray(index) // max(1, shard_size(nvars, world)), world - 1)


def owner_plan(i0, i1, nvars0: int, nvars1: int | None, world: int):
    """Plan of an owner-sharded sweep for one factor family.

    A factor lives on the owner of its LAST variable (its forward proposal is consumed where it is produced); a factor
    whose FIRST variable is owned by another rank is a cut edge: its backward proposal row has to reach that owner, and
    the first variable's particles have to be refreshed on the factor's rank after every product (halo).
    i0 / i1: variable index of every factor's first / last variable (i1 = None for priors, which live on owner(i0)).
    Returns a dict:
      order        factor ids sorted by rank (stable): upload the table in this order so every rank's share is a range
      ranges       [(first, count)] per rank into `order`
      cut          boolean per factor (original numbering)
      send_bwd     {(src_rank, dst_rank): factor ids} backward rows that cross
      halo         [sorted unique first-variable indices read on rank r but owned elsewhere] per rank"""
    import numpy as np
    i0 = np.asarray(i0, dtype=np.int64)
    o0 = owner_of(i0, nvars0, world)
    if i1 is None:
        rank_of, cut = o0, np.zeros(len(i0), bool)
    else:
        rank_of = owner_of(np.asarray(i1, dtype=np.int64), nvars1 if nvars1 is not None else nvars0, world)
        cut = rank_of != o0
    order = np.argsort(rank_of, kind="stable")
    counts = np.bincount(rank_of, minlength=world)
    firsts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    send = {}
    for f in np.nonzero(cut)[0]:
        send.setdefault((int(rank_of[f]), int(o0[f])), []).append(int(f))
    halo = [np.unique(i0[cut & (rank_of == r)]) for r in range(world)]
    return dict(order=order, ranges=[(int(a), int(c)) for a, c in zip(firsts, counts)], cut=cut,
                send_bwd={k: np.asarray(v, dtype=np.int64) for k, v in send.items()}, halo=halo, rank_of=rank_of)


def exchange_bytes(plan, n_factors: int, world: int, row_bytes: int, block_bytes: int, directions: int = 2):
    """bytes RECEIVED per rank and sweep: (all-gather of every proposal row, as built today) vs (owner exchange: the
    backward rows of cut edges + the refreshed particle blocks of halo variables); worst rank of each"""
    allgather = directions * (n_factors - min(c for _, c in plan["ranges"])) * row_bytes
    recv_rows = [0] * world
    for (src, dst), ids in plan["send_bwd"].items():
        recv_rows[dst] += len(ids)
    owner = max(r * row_bytes + len(h) * block_bytes for r, h in zip(recv_rows, plan["halo"]))
    return allgather, owner
