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
