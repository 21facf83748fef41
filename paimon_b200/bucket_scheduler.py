"""Bucket -> GPU scheduling for the multi-GPU merge path.

A Paimon split is one (partition, bucket) (reference: MergeFileSplitRead.createReader,
paimon-core/.../operation/MergeFileSplitRead.java:231-247) and a key lives in exactly one bucket
(bucket/DefaultBucketFunction.java:31-34), so buckets merge independently: every rank owns a subset of the
buckets and there is NO data-path collective.  The only cross-rank traffic is the metrics reduction below
(rows merged, max step time), which is what the reference's CompactionMetrics.Reporter aggregates per task.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence


def assign_buckets(n_buckets: int, world_size: int, weights: Optional[Sequence[float]] = None) -> List[List[int]]:
    """Buckets per rank.  Without weights: bucket b -> rank b mod G.  With weights (e.g. sum of file sizes,
    DataFileMeta.fileSize): longest-processing-time-first, ties by bucket id, deterministic on every rank."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    out: List[List[int]] = [[] for _ in range(world_size)]
    if weights is None:
        for b in range(n_buckets):
            out[b % world_size].append(b)
        return out
    if len(weights) != n_buckets:
        raise ValueError("one weight per bucket")
    load = [0.0] * world_size
    for b in sorted(range(n_buckets), key=lambda i: (-float(weights[i]), i)):
        r = min(range(world_size), key=lambda i: (load[i], i))
        out[r].append(b)
        load[r] += float(weights[b])
    for lst in out:
        lst.sort()
    return out


def my_buckets(rank: int, n_buckets: int, world_size: int, weights: Optional[Sequence[float]] = None) -> List[int]:
    return assign_buckets(n_buckets, world_size, weights)[rank]


def reduce_stats(local: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """Whole-job view of per-rank counters: keys ending in '_max' / '_ms' are MAX-reduced (time is the slowest
    rank's), everything else is summed.  Works on gloo (CPU tests) and nccl."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    keys = sorted(local)
    sums = torch.tensor([float(local[k]) for k in keys], dtype=torch.float64, device=device)
    maxs = sums.clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    return {k: float(maxs[i] if (k.endswith("_max") or k.endswith("_ms")) else sums[i]) for i, k in enumerate(keys)}
