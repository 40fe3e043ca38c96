"""
Multi-GPU evaluation: the walker batch is sharded across the ranks of a ``torch.distributed`` group
(one process per GPU, NCCL over NVLink; gloo on CPU for tests).  Walkers are independent
(``Posterior.log_like`` is a pure function of one point; the reference farms points to workers the same
way, bajes/pipe/utils/mpi.py:120-159), so the data path has NO collective: each walker's whole reduction
happens on one GPU and results are bit-identical for any number of ranks.  The only communication is
the plumbing around it -- broadcast of the [W, ndim] parameter block from the sampler rank and an
all-gather of the W log-likelihoods (8 MB at W = 2^20: latency-, not bandwidth-bound).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np


def shard_bounds(n: int, world: int):
    """Contiguous, balanced slices: rank r owns [lo[r], hi[r])."""
    base, rem = divmod(n, world)
    lo = [r * base + min(r, rem) for r in range(world)]
    hi = [lo[r] + base + (1 if r < rem else 0) for r in range(world)]
    return lo, hi


class ShardedEvaluator:
    """``evaluate(X)``: rank ``src`` supplies X [W, ndim]; every rank returns the full logL [W].

    ``local_eval(X_slice) -> logL_slice`` is the per-rank worker: on a GPU box it is
    ``lambda Xs: like.log_like_batch(Xs, names, const)`` (or the device-pointer variant below); the CPU
    tests pass any pure function.
    """

    def __init__(self, local_eval: Callable[[np.ndarray], np.ndarray], ndim: int, group=None, device=None, src=0,
                 local_eval_device=None):
        """``local_eval_device(x_dev [w, ndim], out_dev [w])`` (optional, CUDA): evaluates a device-resident slice in
        place on the current stream -- the shard then never leaves HBM between the broadcast and the all-gather."""
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist
        self.local_eval = local_eval
        self.ndim = int(ndim)
        self.group = group
        self.src = src
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cpu")
        self.local_eval_device = local_eval_device
        self._pin_x = None          # pinned staging of the sampler rank (CUDA only): X in, gathered logL out
        self._pin_out = None

    def _pinned(self, name, numel):
        """A persistent pinned host buffer of at least ``numel`` doubles (CUDA devices only; None on CPU)."""
        if self.device.type != "cuda":
            return None
        buf = getattr(self, name)
        if buf is None or buf.numel() < numel:
            buf = self.torch.empty(max(numel, 1), dtype=self.torch.float64).pin_memory()
            setattr(self, name, buf)
        return buf

    def evaluate(self, X: Optional[np.ndarray]) -> np.ndarray:
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return np.asarray(self.local_eval(np.ascontiguousarray(X, dtype=np.float64)))
        # 1. batch size, then the parameter block, from the sampler rank
        n = torch.zeros(1, dtype=torch.int64, device=self.device)
        if self.rank == self.src:
            n[0] = -1 if X is None else len(X)
        dist.broadcast(n, src=self.src, group=self.group)
        W = int(n.item())
        if W < 0:                       # shutdown signal for the worker loop
            return None
        if self.rank == self.src:
            xh = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64))
            pin = self._pinned("_pin_x", xh.numel())
            if pin is not None:                      # pageable -> pinned -> device: the H2D copy runs at PCIe speed
                stage = pin[:xh.numel()].view(xh.shape)
                stage.copy_(xh)
                xt = stage.to(self.device, non_blocking=True)
            else:
                xt = xh.to(self.device)
        else:
            xt = torch.empty((W, self.ndim), dtype=torch.float64, device=self.device)
        dist.broadcast(xt, src=self.src, group=self.group)
        # 2. each rank evaluates its contiguous slice (no data-path collective)
        lo, hi = shard_bounds(W, self.world)
        mine = xt[lo[self.rank]:hi[self.rank]]
        # 3. all-gather (padded to equal chunks so one collective suffices)
        chunk = max(h - l for l, h in zip(lo, hi))
        send = torch.full((chunk,), float("nan"), dtype=torch.float64, device=self.device)
        if self.local_eval_device is not None and mine.shape[0] > 0:
            self.local_eval_device(mine.contiguous(), send[:mine.shape[0]])
        else:
            out = np.asarray(self.local_eval(mine.cpu().numpy()), dtype=np.float64)
            send[:len(out)] = torch.as_tensor(out).to(self.device)
        recv = torch.empty(chunk * self.world, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        pin = self._pinned("_pin_out", recv.numel()) if self.rank == self.src else None
        if pin is not None:
            host = pin[:recv.numel()]
            host.copy_(recv, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            recv = host.numpy().reshape(self.world, chunk)
        else:
            recv = recv.cpu().numpy().reshape(self.world, chunk)
        return np.concatenate([recv[r, :hi[r] - lo[r]] for r in range(self.world)])


    def serve(self):
        """Worker loop of the non-sampler ranks (the role of MPIPool.wait, bajes/pipe/utils/mpi.py:91-117):
        take part in every evaluate() of the sampler rank until it calls shutdown()."""
        assert self.rank != self.src
        n_calls = 0
        while self.evaluate(None) is not None:
            n_calls += 1
        return n_calls

    def shutdown(self):
        """Sampler rank: release the workers."""
        if self.world > 1 and self.rank == self.src:
            self.evaluate(None)


class FrequencySplitEvaluator:
    """Small batches on several GPUs: instead of sharding walkers (fewer than one 128-walker tile per GPU leaves SMs
    idle), every rank evaluates ALL walkers on its contiguous share of the frequency chunks and the per-walker
    partial sums (2 x n_ifo doubles per walker) are combined with ONE NCCL all-reduce -- the only place on this path
    with a real exchange step.  The summation order differs from the single-GPU tree, so
    results agree to rounding (~1e-13 relative) rather than bit for bit; use it only when W < world x 128."""

    def __init__(self, like, names: Sequence[str], const: dict, group=None, src=0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.like = like
        self.pmap = like.param_map(names, const)
        self.ndim = len(names)
        self.group = group
        self.src = src
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", like.device)
        self.n_ifo = len(like.ifos)
        n_chunks = like.engine.num_chunks()
        lo, hi = shard_bounds(n_chunks, self.world)
        self.chunk_range = (lo[self.rank], hi[self.rank])

    def evaluate(self, X: Optional[np.ndarray]):
        torch, dist = self.torch, self.dist
        n = torch.zeros(1, dtype=torch.int64, device=self.device)
        if self.rank == self.src:
            n[0] = -1 if X is None else len(X)
        if self.world > 1:
            dist.broadcast(n, src=self.src, group=self.group)
        W = int(n.item())
        if W < 0:
            return None
        if self.rank == self.src:
            xt = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64)).to(self.device)
        else:
            xt = torch.empty((W, self.ndim), dtype=torch.float64, device=self.device)
        if self.world > 1:
            dist.broadcast(xt, src=self.src, group=self.group)
        st = torch.cuda.current_stream(self.device).cuda_stream
        sums = torch.empty((2 * self.n_ifo, W), dtype=torch.float64, device=self.device)
        self.like.engine.partial_device(xt.data_ptr(), W, self.ndim, self.pmap, self.chunk_range[0],
                                        self.chunk_range[1], sums.data_ptr(), st)
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        out = torch.empty(W, dtype=torch.float64, device=self.device)
        self.like.engine.finish_device(sums.data_ptr(), W, out.data_ptr(), 0, st)
        return out.cpu().numpy()

    def serve(self):
        n_calls = 0
        while self.evaluate(None) is not None:
            n_calls += 1
        return n_calls

    def shutdown(self):
        if self.world > 1 and self.rank == self.src:
            self.evaluate(None)


class ShardedLikelihood:
    """Likelihood facade for the sampler rank: ``log_like_batch`` shards the batch over all ranks."""

    def __init__(self, like, evaluator):
        self.like = like
        self.evaluator = evaluator
        for attr in ("ifos", "dets", "logZ_noise"):
            if hasattr(like, attr):
                setattr(self, attr, getattr(like, attr))

    def log_like_batch(self, X, names, const=None):
        return self.evaluator.evaluate(np.ascontiguousarray(X, dtype=np.float64))

    def log_like(self, params):
        return self.like.log_like(params)


class DeviceShardedLikelihood:
    """Device-resident variant used by bench.py: X stays in HBM, the kernel entry takes device pointers,
    logL is all-gathered with NCCL.  ``like`` is a ``GWLikelihood`` living on this rank's GPU."""

    def __init__(self, like, names: Sequence[str], const: dict, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.like = like
        self.pmap = like.param_map(names, const)
        self.ndim = len(names)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", like.device)

    def evaluate_local(self, x_dev, out_dev, parts_dev=None):
        """x_dev [w, ndim] float64 CUDA tensor (this rank's slice) -> out_dev [w]; asynchronous on the
        current torch stream."""
        torch = self.torch
        st = torch.cuda.current_stream(self.device).cuda_stream
        self.like.engine.loglike_device(x_dev.data_ptr(), x_dev.shape[0], self.ndim, self.pmap, out_dev.data_ptr(),
                                        parts_dev.data_ptr() if parts_dev is not None else 0, st)
        return out_dev

    def gather(self, out_dev):
        """all-gather of equally sized per-rank logL vectors -> [world * w]."""
        if self.world == 1:
            return out_dev
        recv = self.torch.empty(out_dev.numel() * self.world, dtype=out_dev.dtype, device=out_dev.device)
        self.dist.all_gather_into_tensor(recv, out_dev, group=self.group)
        return recv
