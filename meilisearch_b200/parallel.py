"""Multi-GPU plumbing.

Keyword path — query-sharded ("replicas only"): every rank holds a full replica of the staged index and serves its own slice of
the query stream; no collective on the data path; torch.distributed only agrees on timings and gathers results
when one caller wants the whole batch back.

Vector stage — optionally corpus-sharded by contiguous docid range (cfg 5): every rank scans its rows for the whole
query batch and the per-shard top-k lists are exchanged with ONE all-gather (NCCL on GPUs, gloo in the CPU tests) and merged."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a batch owned by `rank` (balanced to within one item)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_queries(queries, rank: int, world: int):
    lo, hi = shard_bounds(len(queries), rank, world)
    return queries[lo:hi]


def max_over_ranks(value: float, device="cpu") -> float:
    """The number every multi-GPU measurement reports: the slowest rank's time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_hits(local_ids, n_total: int, limit: int, rank: int, world: int, device="cpu"):
    """All-gather per-rank top-`limit` docid rows (padded with 0xffffffff) back into batch order."""
    lo, hi = shard_bounds(n_total, rank, world)
    width = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
    buf = torch.full((width, limit), 0xFFFFFFFF, dtype=torch.int64, device=device)
    if hi > lo:
        buf[: hi - lo] = torch.as_tensor(local_ids, dtype=torch.int64, device=device)
    if world == 1 or not dist.is_initialized():
        return buf[: hi - lo]
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    rows = []
    for r in range(world):
        a, b = shard_bounds(n_total, r, world)
        rows.append(out[r][: b - a])
    return torch.cat(rows, 0)


# ---- corpus-sharded vector stage (cfg 5): shard g owns a contiguous docid range of the embedding matrix ----------
def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous row range [lo, hi) of the embedding matrix owned by `rank` (rows are in docid order)."""
    return shard_bounds(n_rows, rank, world)


def merge_sharded_topk(local_ids, local_dist, local_count, limit: int, device="cpu"):
    """The one exchange step of the corpus-sharded path: every rank scanned ITS rows for the WHOLE query batch and holds, per
    query, its local top-`limit` (ids, distances ascending, count).  One all-gather of B x limit x (u32 docid, f32 distance) per
    rank, then a k-way merge by (distance, docid) — identical to the single-shard result because every global top-k entry is in
    its shard's local top-k.  Returns (ids [B, limit] int64, dist [B, limit] float32, count [B] int64) on every rank."""
    ids = torch.as_tensor(local_ids, device=device).to(torch.int64)
    dst = torch.as_tensor(local_dist, device=device).to(torch.float32)
    cnt = torch.as_tensor(local_count, device=device).to(torch.int64)
    B = ids.shape[0]
    col = torch.arange(limit, device=device)[None, :]
    valid = col < cnt[:, None]
    ids = torch.where(valid, ids, torch.full_like(ids, 0xFFFFFFFF))
    dst = torch.where(valid, dst, torch.full_like(dst, float("inf")))
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world > 1:
        g_ids = [torch.empty_like(ids) for _ in range(world)]
        g_dst = [torch.empty_like(dst) for _ in range(world)]
        dist.all_gather(g_ids, ids)
        dist.all_gather(g_dst, dst)
        ids, dst = torch.cat(g_ids, 1), torch.cat(g_dst, 1)
    # order by (distance, docid): sort by docid first, then a stable sort by distance
    o1 = torch.argsort(ids, dim=1, stable=True)
    ids, dst = torch.gather(ids, 1, o1), torch.gather(dst, 1, o1)
    o2 = torch.argsort(dst, dim=1, stable=True)
    ids, dst = torch.gather(ids, 1, o2)[:, :limit], torch.gather(dst, 1, o2)[:, :limit]
    count = torch.isfinite(dst).sum(1)
    return ids, dst, count
