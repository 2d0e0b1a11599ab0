"""Multi-GPU plumbing: one process per GPU, reads sharded contiguously.

The record stream is cut into contiguous shards (rank g owns global record
indices [lo_g, hi_g)); every rank runs the whole hot loop on its shard with its
own collectors; afterwards the tables are merged with ``torch.distributed``
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  No collective sits
on the data path.

What merges exactly:
  * additive count tables -- QCMetrics, AdapterCounter, InsertSizeMetrics
    histogram, all counters: all-reduce(SUM) after padding to the longest
    ``max_length`` (all-reduce(MAX));
  * NanoStats: per-read records concatenated in rank order, min/max times by
    all-reduce;
  * PerTileQuality: tiles seen by one rank only are taken as they are.  A tile
    whose reads straddle a shard border has order-dependent double sums: it is
    reported in ``straddling`` and merged by adding the partial sums (NOT the
    reference's rounding order) -- shard at tile borders to avoid it.
DedupEstimator / OverrepresentedSequences tables are order dependent across
shards (escalation point, table-full admission) and stay per rank for now.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_records: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, near-equal shards: rank g owns [lo, hi)."""
    base, extra = divmod(n_records, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return None, None, None
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
        else torch.device("cpu")
    return torch, dist, dev


def allreduce_max(value: int) -> int:
    torch, dist, dev = _dist()
    if dist is None:
        return value
    t = torch.tensor([value], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def allreduce_sum_tables(tables: list[np.ndarray]) -> list[np.ndarray]:
    """One all-reduce(SUM) over a list of equally-shaped-per-rank u64 tables."""
    torch, dist, dev = _dist()
    if dist is None:
        return [np.asarray(t, dtype=np.uint64) for t in tables]
    flat = np.concatenate([np.asarray(t, dtype=np.uint64).ravel() for t in tables]) if tables \
        else np.zeros(0, np.uint64)
    t = torch.from_numpy(flat.view(np.int64).copy()).to(dev)
    dist.all_reduce(t)  # two's complement: the wrap-around of int64 equals the u64 sum
    out = t.cpu().numpy().view(np.uint64)
    res, off = [], 0
    for src in tables:
        n = int(np.asarray(src).size)
        res.append(out[off:off + n].reshape(np.asarray(src).shape).copy())
        off += n
    return res


def _pad_rows(table: np.ndarray, rows: int, width: int, anchor_end: bool = False) -> np.ndarray:
    t = np.asarray(table, dtype=np.uint64).reshape(-1, width)
    if t.shape[0] == rows:
        return t
    out = np.zeros((rows, width), dtype=np.uint64)
    if anchor_end:
        out[rows - t.shape[0]:] = t
    else:
        out[:t.shape[0]] = t
    return out


def merge_qc(base, phred, ea_base, ea_phred, gc, phred_scores) -> dict:
    """QCMetrics tables of all ranks summed (per-position tables padded to the longest read)."""
    rows = allreduce_max(len(base) // 5)
    b, p = _pad_rows(base, rows, 5), _pad_rows(phred, rows, 12)
    merged = allreduce_sum_tables([b, p, np.asarray(ea_base, np.uint64), np.asarray(ea_phred, np.uint64),
                                   np.asarray(gc, np.uint64), np.asarray(phred_scores, np.uint64)])
    keys = ("base_count_table", "phred_count_table", "end_anchored_base_count_table",
            "end_anchored_phred_count_table", "gc_content", "phred_scores")
    return {k: v.ravel() for k, v in zip(keys, merged)}


def merge_adapter_counts(counts) -> list:
    """[(adapter, forward, reverse)] summed over ranks."""
    rows = allreduce_max(max([len(f) for _, f, _ in counts], default=0))
    padded = []
    for _, f, r in counts:
        padded += [_pad_rows(f, rows, 1).ravel(), _pad_rows(r, rows, 1).ravel()]
    merged = allreduce_sum_tables(padded)
    return [(a, merged[2 * i], merged[2 * i + 1]) for i, (a, _, _) in enumerate(counts)]


def gather_nanostats(infos: np.ndarray) -> np.ndarray:
    """Per-read NanoStats records of all ranks, concatenated in rank (= read) order."""
    torch, dist, dev = _dist()
    if dist is None:
        return infos
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, infos.tobytes())
    return np.concatenate([np.frombuffer(p, dtype=infos.dtype) for p in parts])


def merge_tile_counts(tiles) -> tuple[list, list]:
    """PerTileQuality.get_tile_counts() of all ranks -> (merged, straddling tile ids).

    Tiles seen by a single rank are exact.  For a tile seen by several ranks the
    partial sums are added in rank order, which is not the reference's rounding
    order; such tiles are returned in ``straddling`` so that the caller can shard
    at tile borders instead."""
    torch, dist, dev = _dist()
    if dist is None:
        return list(tiles), []
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, [(int(t), list(e), list(c)) for t, e, c in tiles])
    merged, straddling = {}, []
    for part in parts:  # rank order = read order
        for t, e, c in part:
            if t not in merged:
                merged[t] = (list(e), list(c))
                continue
            straddling.append(t)
            me, mc = merged[t]
            n = max(len(me), len(e))
            me += [0.0] * (n - len(me))
            mc += [0] * (n - len(mc))
            for i, (x, y) in enumerate(zip(e, c)):
                me[i] += x
                mc[i] += y
    return [(t, merged[t][0], merged[t][1]) for t in sorted(merged)], sorted(set(straddling))
