"""world_size-2 gloo test of the query-sharded multi-GPU plumbing (host logic only; no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, limit, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from meilisearch_b200.parallel import gather_hits, max_over_ranks, shard_bounds, shard_queries

    queries = [f"q{i}" for i in range(n_total)]
    mine = shard_queries(queries, rank, world)
    lo, hi = shard_bounds(n_total, rank, world)
    assert mine == queries[lo:hi]
    local = np.arange(lo, hi)[:, None] * 100 + np.arange(limit)[None, :]
    allhits = gather_hits(local, n_total, limit, rank, world)
    t = max_over_ranks(1.0 + rank)
    if rank == 0:
        ret["hits"] = allhits.numpy().copy()
        ret["t"] = t
    dist.destroy_process_group()


def test_query_sharding_world2():
    n_total, limit, world = 11, 4, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, limit, ret), nprocs=world, join=True)
    want = np.arange(n_total)[:, None] * 100 + np.arange(limit)[None, :]
    assert np.array_equal(ret["hits"], want)
    assert ret["t"] == 2.0


def test_shard_bounds_partition():
    from meilisearch_b200.parallel import shard_bounds

    for n in (0, 1, 7, 1024):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
