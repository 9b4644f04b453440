"""Multi-GPU plumbing for the query-sharded ("replicas only") deployment: every rank holds a full replica of the
staged index and serves its own slice of the query stream; there is no collective on the data path.
torch.distributed is used only to agree on timings and to gather results when one caller wants the whole batch back."""
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
