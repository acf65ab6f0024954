"""N>1 host logic on CPU: world_size-2 gloo processes shard a factor list, fill their slice of a globally indexed
row buffer and all-gather it in place (the exchange bench.py --gpus N performs over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_factors, q):
    sys.path.insert(0, ROOT)
    from rome_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = sharding.shard_size(n_factors, world)
        first, count = sharding.shard_range(n_factors, rank, world)
        rows = torch.full((world * c, 8, 3), -1.0)
        # "evaluate" the shard: row f holds f everywhere
        for f in range(first, first + count):
            rows[f] = float(f)
        sharding.allgather_rows(rows, n_factors)
        ok = all(bool((rows[f] == float(f)).all()) for f in range(n_factors))
        q.put((rank, first, count, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_factors", [10, 11])
def test_factor_sharding_allgather_world2(n_factors):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_factors, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[3] for r in res] == [True, True]
    assert res[0][1] == 0 and res[0][2] + res[1][2] == n_factors and res[1][1] == res[0][2]


def test_shard_ranges_cover_everything():
    from rome_b200 import sharding
    for n in (0, 1, 7, 8, 12000, 12001):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                f, c = sharding.shard_range(n, r, w)
                assert c <= sharding.shard_size(n, w)
                seen += list(range(f, f + c))
            assert seen == list(range(n))


def test_owner_plan_covers_and_cuts():
    """owner partition (planned replacement of the all-gather): every factor on exactly one rank, ranges contiguous in
    the permuted order, cut edges = owner(first) != owner(last), halo = their first variables; on the bench graph the
    owner exchange moves two orders of magnitude fewer bytes than the all-gather"""
    import numpy as np
    from rome_b200 import sharding as S
    rng = np.random.default_rng(0)
    V, F, G = 1000, 3000, 4
    i0 = rng.integers(0, V - 1, F)
    i1 = np.where(rng.random(F) < 0.9, i0 + 1, rng.integers(0, V, F))
    plan = S.owner_plan(i0, i1, V, V, G)
    assert sorted(plan["order"].tolist()) == list(range(F))
    assert sum(c for _, c in plan["ranges"]) == F
    for r, (a, c) in enumerate(plan["ranges"]):
        ids = plan["order"][a:a + c]
        assert np.all(S.owner_of(i1[ids], V, G) == r)
    own0, own1 = S.owner_of(i0, V, G), S.owner_of(i1, V, G)
    assert np.array_equal(plan["cut"], own0 != own1)
    assert sum(len(v) for v in plan["send_bwd"].values()) == int(plan["cut"].sum())
    for (src, dst), ids in plan["send_bwd"].items():
        assert src != dst and np.all(own1[ids] == src) and np.all(own0[ids] == dst)
    for r in range(G):
        assert np.all(S.owner_of(plan["halo"][r], V, G) != r)
        assert set(plan["halo"][r]) == set(i0[plan["cut"] & (own1 == r)])
    pri = S.owner_plan(np.array([0, V - 1]), None, V, None, G)
    assert not pri["cut"].any() and [c for _, c in pri["ranges"]] == [1, 0, 0, 1]
    # the bench graph (10 000 Pose2, 11 999 Pose2Pose2, N = 100 -> 1248-B rows, 1296-B particle blocks)
    import rome_b200 as rb
    fg = rb.generateGraph_ManhattanShaped(10000, seed=2, N=100)
    idx = {l: v.index for l, v in fg.variables.items()}
    fs = [f for f in fg.factors.values() if isinstance(f.fnc, rb.Pose2Pose2)]
    a = np.array([idx[f.variableOrderSymbols[0]] for f in fs])
    b = np.array([idx[f.variableOrderSymbols[1]] for f in fs])
    for world, most in ((2, 0.04), (8, 0.10)):
        p = S.owner_plan(a, b, 10000, 10000, world)
        assert 0 < p["cut"].mean() < most
        ag, ow = S.exchange_bytes(p, len(fs), world, 1248, 1296)
        assert ow * 20 < ag


# ---- owner-sharded exchange (bench.py --gpus N, GibbsSolver(distributed="owner")): plan consistency + a two-rank run ----
def _row_of(gid, d=3, n=8):
    """stand-in for a factor's proposal row: a deterministic function of the GLOBAL factor id"""
    return (np.arange(n * d, dtype=np.float32).reshape(n, d) + 1000.0 * gid)


def _owner_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from rome_b200 import workloads as W
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = W.manhattan_arrays(600, seed=2, N=8)
        sh = W.sharding_of(w, world)
        lv = W.local_view(w, sh, rank)
        f = lv["families"][0]
        # "evaluate": interior rows stay local; every cut row is addressed (destination rank, row) -- gather what every
        # rank sends and let each rank apply the rows addressed to it (the GPUs do this with peer stores)
        cut_ids = f["order"][f["cut_first"]:f["cut_first"] + f["n_cut"]]
        sends = [(int(d), int(r), _row_of(int(g))) for d, r, g in zip(f["dst_rank"], f["dst_row"], cut_ids)]
        halo_sends = []
        for vt, pushes in lv["loc"]["push"].items():
            for reader, local_vars, slots in pushes:
                for v, s in zip(local_vars, slots):
                    halo_sends.append((reader, vt, int(s), lv["particles"][vt][v].copy()))
        everything = [None] * world
        dist.all_gather_object(everything, (sends, halo_sends))
        recv = np.full((len(f["recv"]), 8, 3), -1.0, np.float32)
        parts = {vt: p.copy() for vt, p in lv["particles"].items()}
        for src, (rows, halos) in enumerate(everything):
            for d, r, row in rows:
                if d == rank:
                    assert (recv[r] == -1).all()   # every receive row has exactly one writer
                    recv[r] = row
            for reader, vt, slot, block in halos:
                if reader == rank:
                    parts[vt][slot] = block
        ok_rows = all(np.array_equal(recv[k], _row_of(int(g))) for k, g in enumerate(f["recv"]))
        # after the halo push this rank's particle array equals the global one restricted to (owned + halo) variables
        ok_halo = all(np.array_equal(parts[vt], w["particles"][vt][lv["loc"]["var_global"][vt]]) for vt in parts)
        # local indices address the right global variables
        gid = lv["loc"]["var_global"][0]
        F = w["families"][0]
        ok_idx = np.array_equal(gid[f["i0"]], F["i0"][f["order"]]) and np.array_equal(gid[f["i1"]], F["i1"][f["order"]])
        q.put((rank, bool(ok_rows), bool(ok_halo), bool(ok_idx), len(f["i0"]), int(f["n_cut"]), len(f["recv"])))
    finally:
        dist.destroy_process_group()


def test_owner_sharded_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_owner_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] and r[3] for r in res), res
    assert abs(res[0][4] - res[1][4]) <= 1                       # balanced bounds: same number of factors per rank
    assert res[0][5] + res[1][5] == res[0][6] + res[1][6] > 0    # every cut row is received exactly once


def test_owner_sharding_plan_properties():
    """every factor on exactly one rank, cut = owner(first) != owner(last), receive layouts agree on both sides, halo
    pushes cover exactly the halo sets; two variable types (poses + landmarks); exchange volume on the bench graph"""
    from rome_b200 import sharding as S, workloads as W
    rng = np.random.default_rng(1)
    G, V, Lm, F = 4, 500, 40, 900
    i0 = rng.integers(0, V, F)
    i1 = rng.integers(0, Lm, F)
    sh = S.OwnerSharding(G, {0: V, 1: Lm}, {2: (0, 1, i0, i1), 1: (0, None, np.array([0, V - 1]), None)})
    seen, total_cut, total_recv = [], 0, 0
    for r in range(G):
        L = sh.local(r)
        f = L["fam"][2]
        seen += f["order"].tolist()
        lo, hi = L["own"][0]
        assert np.all((i0[f["order"]] >= lo) & (i0[f["order"]] < hi))
        gid1 = L["var_global"][1]
        assert np.array_equal(gid1[f["i1"]], i1[f["order"]])
        cf = f["cut_first"]
        cut_ids = f["order"][cf:cf + f["n_cut"]]
        interior_ids = np.concatenate([f["order"][:cf], f["order"][cf + f["n_cut"]:]])
        assert cf % 24 == 0 or f["n_cut"] == 0
        assert np.all(sh.owner(1, i1[cut_ids]) != r) and np.all(sh.owner(1, i1[interior_ids]) == r)
        for d, row, g in zip(f["dst_rank"], f["dst_row"], cut_ids):
            assert sh.local(int(d))["fam"][2]["recv"][row] == g
        total_cut += f["n_cut"]
        total_recv += len(f["recv"])
        pushed = {}
        for reader, lv_, slots in L["push"][1]:
            for v, s in zip(lv_, slots):
                pushed[(reader, int(s))] = int(v) + L["own"][1][0]
        for (reader, s), g in pushed.items():
            assert sh.local(reader)["var_global"][1][s] == g
    assert sorted(seen) == list(range(F)) and total_cut == total_recv
    for r in range(G):
        L = sh.local(r)
        n_own = L["own"][1][1] - L["own"][1][0]
        got = sorted(s for o in range(G) if o != r for reader, _, slots in sh.local(o)["push"][1] if reader == r for s in slots)
        assert got == list(range(n_own, n_own + len(L["halo"][1])))
    assert [len(sh.local(r)["fam"][1]["order"]) for r in range(G)] == [1, 0, 0, 1]
    # bench graph: per-rank receive volume of the owner exchange vs the all-gather of every row
    w = W.manhattan_arrays(10000)
    for world in (2, 8):
        sh = W.sharding_of(w, world)
        worst = max(sh.exchange_bytes(r, {0: 1248, 1: 1248}, {0: 1296}) for r in range(world))
        assert worst * 15 < (world - 1) / world * 11999 * 1248   # all-gather: every other rank's rows
        counts = [len(sh.local(r)["fam"][0]["order"]) + len(sh.local(r)["fam"][1]["order"]) for r in range(world)]
        assert max(counts) - min(counts) <= 2
