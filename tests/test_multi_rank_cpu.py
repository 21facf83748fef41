"""world_size-2 gloo test of the N>1 host path: bucket scheduling + metrics reduction (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from paimon_b200.bucket_scheduler import assign_buckets, my_buckets, reduce_stats


def test_assignment_is_a_partition():
    for n, g in [(64, 8), (7, 2), (3, 8), (0, 4)]:
        parts = assign_buckets(n, g)
        assert sorted(b for p in parts for b in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    w = [5, 1, 1, 1, 4, 4, 3, 9]
    parts = assign_buckets(8, 3, w)
    assert sorted(b for p in parts for b in p) == list(range(8))
    loads = [sum(w[b] for b in p) for p in parts]
    assert max(loads) <= 10          # LPT: optimum is ceil(28/3) = 10
    assert parts == assign_buckets(8, 3, w)      # deterministic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = my_buckets(rank, 11, world, weights=[b + 1 for b in range(11)])
        rows = sum(1000 * (b + 1) for b in mine)
        stats = reduce_stats({"rows_in": rows, "buckets": len(mine), "step_ms": 10.0 + rank})
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        q.put((rank, stats, gathered))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, stats, gathered in results:
        assert sorted(b for g in gathered for b in g) == list(range(11))     # every bucket merged exactly once
        assert stats["rows_in"] == sum(1000 * (b + 1) for b in range(11))   # whole-job rows
        assert stats["buckets"] == 11
        assert stats["step_ms"] == 11.0                                      # max over ranks
