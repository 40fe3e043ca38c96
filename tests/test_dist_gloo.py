"""
world_size-2 gloo test (CPU) of the multi-GPU host logic: one scatter of the parameter block (header + rows per rank),
contiguous walker shards, results back in input order; N-rank result == 1-rank result bit for bit.
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


def _worker(rank, world, port, n, q, capacity, device_fn):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from bajes_b200.dist import ShardedEvaluator
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        def eval_dev(x, out):          # the device-pointer worker signature, on CPU tensors
            out.copy_(torch.as_tensor(_fake_logl(x.numpy())))

        ev = ShardedEvaluator(_fake_logl, ndim=5, capacity=capacity, local_eval_device=eval_dev if device_fn else None)
        if rank == 0:                  # only the sampler rank holds the points
            rng = np.random.default_rng(0)
            X = rng.normal(size=(n, 5))
            out = ev.evaluate(X)
            ev.shutdown()
            q.put((rank, out))
        else:
            q.put((rank, ev.serve()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,capacity,device_fn", [(1, 4096, False), (7, 4096, True), (64, 64, False), (1001, 8, False),
                                                  (1001, 2048, True)])
def test_sharded_evaluator_world2(n, capacity, device_fn):
    """one scatter out (header + rows per rank), one gather back; capacity grows collectively when a batch exceeds it"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q, capacity, device_fn)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = np.random.default_rng(0).normal(size=(n, 5))
    ref = _fake_logl(X)
    assert np.array_equal(res[0], ref) and res[1] == 1


def test_shard_bounds_cover_and_balance():
    from bajes_b200.dist import shard_bounds
    for n in (0, 1, 5, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            lo, hi = shard_bounds(n, world)
            assert lo[0] == 0 and hi[-1] == n
            assert all(hi[r] == lo[r + 1] for r in range(world - 1))
            sizes = [h - l for l, h in zip(lo, hi)]
            assert max(sizes) - min(sizes) <= 1


def _serve_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bajes_b200.dist import ShardedEvaluator
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        ev = ShardedEvaluator(_fake_logl, ndim=5)
        if rank == 0:
            rng = np.random.default_rng(1)
            outs = []
            for n in (9, 2, 33):
                X = rng.normal(size=(n, 5))
                outs.append(np.array_equal(ev.evaluate(X), _fake_logl(X)))
            ev.shutdown()
            q.put((rank, outs))
        else:
            q.put((rank, ev.serve()))
    finally:
        dist.destroy_process_group()


def test_worker_loop_serve_and_shutdown():
    """Rank 0 plays the sampler, rank 1 sits in serve() until shutdown (the MPIPool master/worker pattern)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_serve_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == [True, True, True] and res[1] == 3


def test_ensemble_sampler_through_batched_pool():
    """The stretch-move sampler only sees pool.map(posterior.log_post, pts): with BatchedPool that is 2 batched calls
    per step; a Gaussian toy posterior is recovered."""
    from bajes_b200.ensemble import BoxPrior, EnsembleSampler
    from bajes_b200.likelihood import Posterior
    from bajes_b200.pool import BatchedPool

    class ToyLike:
        logZ_noise = 0.0

        def log_like(self, p):
            return -0.5 * ((p["a"] - 1.0) ** 2 / 0.04 + (p["b"] + 2.0) ** 2 / 0.25)

        def log_like_batch(self, X, names, const=None):
            a, b = X[:, names.index("a")], X[:, names.index("b")]
            return -0.5 * ((a - 1.0) ** 2 / 0.04 + (b + 2.0) ** 2 / 0.25)

    prior = BoxPrior(["a", "b"], [[-5, 5], [-5, 5]], {"c": 3.0})
    post = Posterior(ToyLike(), prior)
    pool = BatchedPool(post)
    s = EnsembleSampler(post, 64, pool, seed=3)
    rng = np.random.default_rng(0)
    s.initialize(prior.prior_transform(rng.random((64, 2))))
    s.run(300)
    assert pool.n_batched_calls == 1 + 2 * 300
    s.pos_hist = []
    chain = []
    for _ in range(200):
        s.step()
        chain.append(s.pos.copy())
    chain = np.concatenate(chain)
    assert abs(chain[:, 0].mean() - 1.0) < 0.05 and abs(chain[:, 1].mean() + 2.0) < 0.1
    assert abs(chain[:, 0].std() - 0.2) < 0.05 and abs(chain[:, 1].std() - 0.5) < 0.1
    # batched and per-point posteriors agree row by row
    X = prior.prior_transform(rng.random((5, 2)))
    assert np.allclose(post.log_post_batch(X), [post.log_post(x) for x in X])
