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


# ---- owner-sharded sweeps: variables AND factors partitioned, only cut edges cross ------------------------------------
def balanced_bounds(i0_lists, nvars: int, world: int):
    """contiguous variable ranges [b[r], b[r+1]) such that the factors placed by their first variable (i0 of every
    family given) are spread evenly over the ranks (a rank's work is its factor count, not its variable count)"""
    import numpy as np
    load = np.zeros(nvars, dtype=np.int64)
    for i0 in i0_lists:
        load += np.bincount(np.asarray(i0, dtype=np.int64), minlength=nvars)
    cum = np.concatenate([[0], np.cumsum(load)])
    total = cum[-1]
    b = [0]
    for r in range(1, world):
        b.append(int(np.searchsorted(cum, total * r / world, side="left")))
    b.append(nvars)
    b = np.maximum.accumulate(np.asarray(b, dtype=np.int64))
    return b


def cut_block_start(n_interior: int, n_cut: int, ft: int = 0, grid: int = 0) -> int:
    """local index at which a rank's cut block starts inside its factor table: behind ~40 % of the interior factors,
    tile aligned (24 = lcm of the 8- and 12-factor tiles).  When the launch geometry is known (`ft` factors per tile, `grid`
    persistent CTAs) the start is moved to the nearest tile whose CTAs do NOT also take one of the grid's leftover tiles:
    with T tiles on `grid` CTAs the first T mod grid CTAs run one tile more than the others, and a cut tile is heavier
    than an interior one (it also writes the forward row into peer memory) -- both on one CTA set the kernel's tail.
    ROME_B200_CUT_PLACEMENT=legacy keeps the plain 40 % rule (A/B measurements)."""
    import os
    base = (int(0.4 * n_interior) // 24) * 24 if n_cut else n_interior
    if not n_cut or ft <= 0 or grid <= 0 or os.environ.get("ROME_B200_CUT_PLACEMENT") == "legacy":
        return base
    tiles = -(-(n_interior + n_cut) // ft)
    tail, nc = tiles % grid, -(-n_cut // ft) + 1
    if tiles <= grid or tail == 0 or nc > grid - tail:
        return base
    step = 24 // ft if 24 % ft == 0 else 24       # tile indices whose first factor is a multiple of 24
    best = None
    for r in range(tiles // grid + 1):
        lo, hi = r * grid + tail, r * grid + grid - nc
        c = min(max(base // ft, lo), hi)
        c = -(-c // step) * step
        if lo <= c <= hi and c * ft <= n_interior and (best is None or abs(c * ft - base) < abs(best - base)):
            best = c * ft
    return base if best is None else best


class OwnerSharding:
    """Plan of an owner-sharded sweep.

    Every variable type is split into contiguous ranges, one per rank (`bounds[vt]`, world + 1 entries).  A factor is
    evaluated on the owner of its FIRST variable; if its LAST variable belongs to another rank it is a CUT factor: that
    rank's copy of the last variable is a HALO block (pushed by the owner after every belief update) and the factor's
    forward proposal row is written straight into a receive buffer of the owner (rome_b200_set_proposal_destinations).
    Everything is derived from the global index arrays, identically on every rank, so no plan has to be exchanged.

    families: {family id: (vt0, vt1 or None, i0 array, i1 array or None)}    nvars: {vt: count}"""

    def __init__(self, world: int, nvars: dict, families: dict, bounds: dict | None = None, geometry: dict | None = None):
        import numpy as np
        self.world, self.nvars, self.families = world, dict(nvars), {}
        self.geometry = dict(geometry or {})   # {family: (factors per tile, persistent CTAs)} of the evaluation launches
        self.bounds = {}
        for vt, n in nvars.items():
            if bounds and vt in bounds:
                b = np.asarray(bounds[vt], dtype=np.int64)
            else:
                c = shard_size(n, world)
                b = np.minimum(np.arange(world + 1, dtype=np.int64) * c, n)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
            self.bounds[vt] = b
        for fam, (vt0, vt1, i0, i1) in families.items():
            i0 = np.asarray(i0, dtype=np.int64)
            r0 = self.owner(vt0, i0)
            if vt1 is None or i1 is None:
                i1, r1 = None, r0
            else:
                i1 = np.asarray(i1, dtype=np.int64)
                r1 = self.owner(vt1, i1)
            self.families[fam] = dict(vt0=vt0, vt1=vt1 if i1 is not None else None, i0=i0, i1=i1, rank=r0, dst=r1)
        self._local = {}

    def owner(self, vt, index):
        import numpy as np
        return np.searchsorted(self.bounds[vt], np.asarray(index), side="right") - 1

    def recv_layout(self, fam, rank):
        """global ids of the cut factors whose forward rows land on `rank`, in receive-buffer row order (by source
        rank, then by global id), and the row offset of every source rank"""
        import numpy as np
        F = self.families[fam]
        ids = np.nonzero((F["dst"] == rank) & (F["rank"] != rank))[0]
        order = np.lexsort((ids, F["rank"][ids]))
        return ids[order]

    def local(self, rank: int):
        """everything rank `rank` needs: its variables (owned, then halo), its factors with local indices -- in table
        order [interior part | cut factors (local indices cut_first .. cut_first + n_cut) | rest of the interior] --,
        the destination (rank, row) of every cut factor's forward row, and the halo blocks it must push"""
        import numpy as np
        if rank in self._local:
            return self._local[rank]
        own = {vt: (int(b[rank]), int(b[rank + 1])) for vt, b in self.bounds.items()}
        halo = {vt: [] for vt in self.bounds}
        for fam, F in self.families.items():
            if F["i1"] is None:
                continue
            cut = (F["rank"] == rank) & (F["dst"] != rank)
            halo[F["vt1"]].append(F["i1"][cut])
        halo = {vt: (np.unique(np.concatenate(v)) if v else np.zeros(0, np.int64)) for vt, v in halo.items()}
        var_global = {vt: np.concatenate([np.arange(*own[vt], dtype=np.int64), halo[vt]]) for vt in self.bounds}

        def to_local(vt, gid):
            lo, hi = own[vt]
            inside = (gid >= lo) & (gid < hi)
            pos = np.searchsorted(halo[vt], gid)
            pos = np.minimum(pos, max(len(halo[vt]) - 1, 0))
            if len(halo[vt]):
                assert (inside | (halo[vt][pos] == gid)).all()
            else:
                assert inside.all()
            return np.where(inside, gid - lo, (hi - lo) + pos).astype(np.int32)

        fams = {}
        for fam, F in self.families.items():
            mine = np.nonzero(F["rank"] == rank)[0]
            is_cut = F["dst"][mine] != rank
            interior, cutf = mine[~is_cut], mine[is_cut]
            cutf = cutf[np.lexsort((cutf, F["dst"][cutf]))]  # grouped by destination rank
            # local table order: the cut block sits BEHIND the first ~40 % of the interior factors (tile aligned): when a
            # persistent grid reaches it, the peers' signals of the previous step have long arrived (the barrier is only
            # passed there), and the NVLink latency of its rows hides behind the interior factors that follow
            # (cut_block_start: the rule, refined by the launch geometry when it is known)
            cut_first = cut_block_start(len(interior), len(cutf), *self.geometry.get(fam, (0, 0)))
            order = np.concatenate([interior[:cut_first], cutf, interior[cut_first:]])
            dst_rank = F["dst"][cutf]
            dst_row = np.zeros(len(cutf), dtype=np.int64)
            for d in np.unique(dst_rank):
                lay = self.recv_layout(fam, int(d))
                sel = dst_rank == d
                pos = {int(g): k for k, g in enumerate(lay)}
                dst_row[sel] = [pos[int(g)] for g in cutf[sel]]
            fams[fam] = dict(order=order, n_interior=len(interior), n_cut=len(cutf), cut_first=cut_first,
                             i0=to_local(F["vt0"], F["i0"][order]),
                             i1=to_local(F["vt1"], F["i1"][order]) if F["i1"] is not None else None,
                             dst_rank=dst_rank, dst_row=dst_row, recv=self.recv_layout(fam, rank))
        # halo blocks this rank owns and must push: (reader rank, local variable id here, slot in the reader's store)
        push = {vt: [] for vt in self.bounds}
        for other in range(self.world):
            if other == rank:
                continue
            oown = {vt: (int(b[other]), int(b[other + 1])) for vt, b in self.bounds.items()}
            for vt in self.bounds:
                need = []
                for fam, F in self.families.items():
                    if F["i1"] is None or F["vt1"] != vt:
                        continue
                    cut = (F["rank"] == other) & (F["dst"] == rank)
                    need.append(F["i1"][cut])
                need = np.unique(np.concatenate(need)) if need else np.zeros(0, np.int64)
                if len(need) == 0:
                    continue
                # the reader's halo list is the sorted unique set of ALL its foreign last variables of this type
                ohalo = []
                for fam, F in self.families.items():
                    if F["i1"] is None or F["vt1"] != vt:
                        continue
                    ohalo.append(F["i1"][(F["rank"] == other) & (F["dst"] != other)])
                ohalo = np.unique(np.concatenate(ohalo))
                slot = (oown[vt][1] - oown[vt][0]) + np.searchsorted(ohalo, need)
                push[vt].append((other, (need - own[vt][0]).astype(np.int32), slot.astype(np.int64)))
        out = dict(own=own, halo=halo, var_global=var_global, fam=fams, push=push)
        self._local[rank] = out
        return out

    def exchange_bytes(self, rank, row_bytes: dict, block_bytes: dict):
        """bytes RECEIVED by `rank` per sweep: forward rows of cut factors targeting it + halo blocks it reads"""
        L = self.local(rank)
        return sum(len(f["recv"]) * row_bytes[fam] for fam, f in L["fam"].items()) + \
            sum(len(h) * block_bytes[vt] for vt, h in L["halo"].items())
