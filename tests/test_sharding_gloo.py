"""CPU, world_size 2 (gloo): the N>1 host path -- shard bounds, ragged shards, gather order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, ndim, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mcalf_b200.distributed import ShardedLikelihood, shard_bounds
        P = np.random.default_rng(7).random((B, ndim))
        calls = []

        def fake(rows):          # stand-in for the CUDA likelihood: any pure per-row function
            calls.append(rows.shape[0])
            return rows[:, 0] * 2.0 - rows[:, ndim - 1]            # (one rounding each: same bits in numpy)

        sl = ShardedLikelihood(evaluate=fake, device=torch.device("cpu"))
        out = sl.lnlhood_batch(P)
        expect = P[:, 0] * 2.0 - P[:, ndim - 1]
        ok = out.shape == (B,) and np.array_equal(out.numpy(), expect)
        lo, hi = shard_bounds(B, world)[rank]
        ok = ok and sl.last_shard == (lo, hi) and (calls == ([hi - lo] if hi > lo else []))
        q.put((rank, bool(ok), sl.last_shard))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [0, 1, 7, 64])
def test_sharded_gather_world2(B):
    world, ndim = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, ndim, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    shards = sorted(s for _, _, s in res)
    assert shards[0][0] == 0 and shards[-1][1] == B
    assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))     # contiguous, no overlap


def test_shard_bounds():
    from mcalf_b200.distributed import shard_bounds
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(3, 8)[3:] == [(3, 3)] * 5
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    for B in (1, 5, 262144, 262145):
        for w in (1, 2, 4, 8):
            b = shard_bounds(B, w)
            assert b[0][0] == 0 and b[-1][1] == B and all(x[1] == y[0] for x, y in zip(b, b[1:]))
