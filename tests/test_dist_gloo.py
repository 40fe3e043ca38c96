"""
world_size-2 gloo test (CPU) of the multi-GPU host logic: broadcast of the parameter block, contiguous
walker shards, all-gather of logL in input order; N-rank result == 1-rank result bit for bit.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_logl(X):
    # any pure per-row function; irrational mixing so that order mistakes cannot cancel
    return np.sin(X).sum(axis=1) + 1e-3 * X[:, 0] * X[:, -1]


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bajes_b200.dist import ShardedEvaluator
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        ev = ShardedEvaluator(_fake_logl, ndim=5)
        rng = np.random.default_rng(0)
        X = rng.normal(size=(n, 5))
        out = ev.evaluate(X if rank == 0 else None)      # only the sampler rank holds the points
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 7, 64, 1001])
def test_sharded_evaluator_world2(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = np.random.default_rng(0).normal(size=(n, 5))
    ref = _fake_logl(X)
    assert np.array_equal(res[0], ref) and np.array_equal(res[1], ref)


def test_shard_bounds_cover_and_balance():
    from bajes_b200.dist import shard_bounds
    for n in (0, 1, 5, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            lo, hi = shard_bounds(n, world)
            assert lo[0] == 0 and hi[-1] == n
            assert all(hi[r] == lo[r + 1] for r in range(world - 1))
            sizes = [h - l for l, h in zip(lo, hi)]
            assert max(sizes) - min(sizes) <= 1
