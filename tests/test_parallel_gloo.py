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


def _vec_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from corpus.pyindexgen import IndexImage
    from meilisearch_b200.parallel import merge_sharded_topk, shard_rows
    from oracle.pyoracle import OracleIndex

    rng = np.random.default_rng(5)
    n, d, B, k = 4000, 32, 9, 10
    emb = rng.standard_normal((n, d)).astype(np.float32)
    emb[77] = emb[3001]  # an exact tie across the two shards
    q = rng.standard_normal((B, d)).astype(np.float32)
    q[0] = emb[77]
    img = IndexImage(1)
    img.add_text(0, 0, "doc")
    img.build()
    lo, hi = shard_rows(n, rank, world)
    shard = OracleIndex(img)   # the oracle stands in for the per-GPU scan in this CPU test
    shard.set_embeddings(emb[lo:hi], np.arange(lo, hi, dtype=np.uint32))
    ids = np.full((B, k), 0xFFFFFFFF, np.int64)
    dst = np.full((B, k), np.inf, np.float32)
    cnt = np.zeros(B, np.int64)
    for i in range(B):
        oi, od = shard.nns(q[i], k)
        ids[i, : len(oi)], dst[i, : len(oi)], cnt[i] = oi, od, len(oi)
    m_ids, m_dst, m_cnt = merge_sharded_topk(ids, dst, cnt, k)
    if rank == 0:
        full = OracleIndex(img)
        full.set_embeddings(emb, np.arange(n, dtype=np.uint32))
        want = [full.nns(q[i], k) for i in range(B)]
        ret["ok"] = all(m_ids[i, : m_cnt[i]].tolist() == want[i][0].tolist() and np.array_equal(m_dst[i, : m_cnt[i]].numpy(), want[i][1])
                        for i in range(B))
    dist.destroy_process_group()


def test_corpus_sharded_vector_topk_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_vec_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["ok"]
