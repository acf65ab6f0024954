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
