"""CPU tests of the multi-GPU host logic: partitioning and the score gather over torch.distributed (gloo, world 2)."""
import os
import socket

import numpy as np
import pytest

from tracy_b200 import shard


def test_partition_covers_everything():
    for n in (0, 1, 7, 100, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard.partition(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_partition_by_cost_balances():
    rng = np.random.default_rng(0)
    l1 = rng.integers(100, 2000, 5000)
    l2 = rng.integers(1000, 8000, 5000)
    b = shard.partition_by_cost(l1, l2, 8)
    assert b[0] == 0 and b[-1] == 5000 and all(x <= y for x, y in zip(b, b[1:]))
    cost = l1.astype(np.int64) * l2
    parts = [cost[b[i]:b[i + 1]].sum() for i in range(8)]
    assert max(parts) < 1.05 * (cost.sum() / 8)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1001
    allscores = torch.arange(n, dtype=torch.int32) * 3 - 500
    counts = [shard.partition(n, world, r)[1] - shard.partition(n, world, r)[0] for r in range(world)]
    lo, hi = shard.partition(n, world, rank)
    ref = torch.zeros(64, dtype=torch.uint8) if rank else torch.arange(64, dtype=torch.uint8)
    shard.broadcast_reference(ref, src=0)
    got = shard.gather_scores(allscores[lo:hi].clone(), counts)
    ok = bool(torch.equal(got, allscores)) and bool(torch.equal(ref, torch.arange(64, dtype=torch.uint8)))
    dist.destroy_process_group()
    q.put((rank, ok))


def test_gather_scores_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
