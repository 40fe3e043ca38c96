"""
Multi-GPU evaluation: the walker batch is sharded across the ranks of a ``torch.distributed`` group
(one process per GPU, NCCL over NVLink; gloo on CPU for tests).  Walkers are independent
(``Posterior.log_like`` is a pure function of one point; the reference farms points to workers the same
way, bajes/pipe/utils/mpi.py:120-159), so the data path has NO collective: each walker's whole reduction
happens on one GPU and results are bit-identical for any number of ranks.  The only communication is
the plumbing around it -- one scatter of the [W, ndim] parameter block from the sampler rank, and the
W log-likelihoods stored by every rank's epilogue kernel straight into the sampler rank's HBM over NVLink
(peer memory; a gather on CPU).
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


class _RawCuda:
    """Wraps a raw device address as a float64 vector for ``torch.as_tensor`` (CUDA array interface)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


HDR = 3                     # per-rank header of the scattered block: rows of this rank, offset of its first row, sequence
CMD_SHUTDOWN, CMD_GROW = -1.0, -2.0


class ShardedEvaluator:
    """``evaluate(X)``: rank ``src`` (the sampler) supplies X [W, ndim] and gets logL [W] back; every rank evaluates a
    contiguous slice.  Per call there is ONE collective and, on the sampler rank, ONE host synchronisation:

      out    the sampler rank packs, per rank, [rows, first row, sequence | that rank's rows] into a fixed-capacity
             block and ONE ``scatter`` (NCCL over NVLink) delivers it -- 1/N of the batch per link instead of a
             broadcast of all of it, and no separate size message;
      back   no collective at all on CUDA: the sampler rank owns the result vector, the other ranks map it through a
             CUDA IPC handle, and their epilogue kernel stores every logL straight into the sampler rank's HBM over
             NVLink (``engine.PeerBuffer``; a system-scope flag per rank, awaited by a one-block kernel on the sampler
             rank's stream, orders the device-to-host copy behind the remote stores).  On CPU (gloo tests) or without
             peer access the slices come back with one ``gather``.

    ``local_eval(X_slice) -> logL_slice`` is the per-rank host worker (CPU tests pass any pure function);
    ``local_eval_device(x_dev [w, ndim], out_dev [w])`` evaluates a device-resident slice in place on the current
    stream -- ``out_dev`` may be a view of the sampler rank's memory.  The capacity grows on demand (collectively).
    The other ranks sit in ``serve()`` (the role of MPIPool.wait, bajes/pipe/utils/mpi.py:91-117)."""

    def __init__(self, local_eval: Optional[Callable[[np.ndarray], np.ndarray]], ndim: int, group=None, device=None,
                 src=0, local_eval_device=None, capacity: int = 4096, peer: Optional[bool] = None):
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
        self.seq = 0
        self.peer = None
        self.want_peer = (self.device.type == "cuda" and local_eval_device is not None) if peer is None else bool(peer)
        self.capacity = 0
        if self.world > 1:
            self._setup(int(capacity))

    # -- collective set-up (every rank, same capacity) ------------------------------------------------------------
    def _setup(self, capacity: int):
        torch, dist = self.torch, self.dist
        self.capacity = int(capacity)
        self.cap_rank = -(-self.capacity // self.world)
        self.slot = HDR + self.cap_rank * self.ndim
        self.recv = torch.empty(self.slot, dtype=torch.float64, device=self.device)
        cuda = self.device.type == "cuda"
        if self.rank == self.src:
            self.stage = torch.empty((self.world, self.slot), dtype=torch.float64, pin_memory=cuda)
            self.send = self.stage if not cuda else torch.empty((self.world, self.slot), dtype=torch.float64,
                                                                 device=self.device)
            self.host_out = torch.empty(2 + self.capacity, dtype=torch.float64, pin_memory=cuda)
        if self.peer is not None:
            self.peer.close()
            self.peer = None
        self.out_view = None
        if self.want_peer:
            from . import engine
            ok = torch.ones(1, dtype=torch.int32, device=self.device)
            handle = [None]
            try:
                if self.rank == self.src:
                    self.peer = engine.PeerBuffer.create(self.device.index, self.capacity)
                    handle[0] = self.peer.handle
            except Exception:
                ok.zero_()
            dist.broadcast_object_list(handle, src=self.src, group=self.group)
            if self.rank != self.src:
                try:
                    self.peer = engine.PeerBuffer.open(self.device.index, handle[0], self.capacity)
                except Exception:
                    ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:                      # some pair of GPUs has no peer access: gather instead
                if self.peer is not None:
                    self.peer.close()
                self.peer = None
            else:
                # [status, pad, logL...] as a tensor (the flags sit in front of it)
                from .engine import PEER_FLAGS
                self.out_all = torch.as_tensor(_RawCuda(self.peer.base + PEER_FLAGS * 8, 2 + self.capacity),
                                               device=self.device)
                self.out_view = self.out_all[2:]
        if self.peer is None:
            self.ret = torch.empty(self.cap_rank, dtype=torch.float64, device=self.device)
            self.ret_all = ([torch.empty(self.cap_rank, dtype=torch.float64, device=self.device)
                             for _ in range(self.world)] if self.rank == self.src else None)

    # -- one batch ----------------------------------------------------------------------------------------------------
    def _scatter(self):
        self.dist.scatter(self.recv, scatter_list=list(self.send.unbind(0)) if self.rank == self.src else None,
                          src=self.src, group=self.group)

    def _command(self, cmd: float, arg: float = 0.0):
        """sampler rank: a header-only block for every rank (shutdown / grow)"""
        self.stage[:, 0] = cmd
        self.stage[:, 1] = arg
        if self.send is not self.stage:
            self.send[:, :HDR].copy_(self.stage[:, :HDR], non_blocking=True)
        self._scatter()

    def _local(self, n: int, lo: int):
        """evaluate rows [lo, lo + n) of the batch, which sit in self.recv, and deliver them"""
        torch = self.torch
        rows = self.recv[HDR:HDR + n * self.ndim].view(n, self.ndim)
        if self.peer is not None:
            if n > 0:
                self.local_eval_device(rows, self.out_view[lo:lo + n])
            if self.rank != self.src:
                self.peer.signal(self.rank, self.seq, torch.cuda.current_stream(self.device).cuda_stream)
            return
        if n > 0:
            if self.local_eval_device is not None:
                self.local_eval_device(rows, self.ret[:n])
            else:
                out = np.asarray(self.local_eval(rows.cpu().numpy()), dtype=np.float64)
                self.ret[:n] = torch.as_tensor(out).to(self.device)
        self.dist.gather(self.ret, gather_list=self.ret_all, dst=self.src, group=self.group)

    def evaluate(self, X: np.ndarray) -> np.ndarray:
        torch = self.torch
        X = np.ascontiguousarray(X, dtype=np.float64)
        if self.world == 1:
            return np.asarray(self.local_eval(X))
        assert self.rank == self.src, "evaluate() is the sampler rank's call; the other ranks run serve()"
        W = len(X)
        if W > self.capacity:
            new_cap = max(W, 2 * self.capacity)
            self._command(CMD_GROW, float(new_cap))
            self._setup(new_cap)
        self.seq += 1
        lo, hi = shard_bounds(W, self.world)
        st = self.stage.numpy()
        for r in range(self.world):
            n = hi[r] - lo[r]
            st[r, 0], st[r, 1], st[r, 2] = n, lo[r], self.seq
            st[r, HDR:HDR + n * self.ndim] = X[lo[r]:hi[r]].reshape(-1)
        if self.send is not self.stage:
            used = HDR + max(h - l for l, h in zip(lo, hi)) * self.ndim
            if used * 4 >= self.slot * 3:
                self.send.copy_(self.stage, non_blocking=True)                    # one H2D copy
            else:
                self.send[:, :used].copy_(self.stage[:, :used], non_blocking=True)
        self._scatter()
        self._local(hi[self.src] - lo[self.src], lo[self.src])
        if self.peer is not None:
            stream = torch.cuda.current_stream(self.device)
            others = [r for r in range(self.world) if r != self.src]
            # flags are indexed by rank; the sampler's own slot is never awaited
            first, last = min(others), max(others)
            if self.src > first and self.src < last:
                self.peer.wait(first, self.src - first, self.seq, stream=stream.cuda_stream)
                self.peer.wait(self.src + 1, last - self.src, self.seq, stream=stream.cuda_stream)
            else:
                self.peer.wait(first, last - first + 1, self.seq, stream=stream.cuda_stream)
            self.host_out[:2 + W].copy_(self.out_all[:2 + W], non_blocking=True)
            stream.synchronize()
            if int(self.host_out[:1].view(torch.int64)[0]) != 0:
                raise RuntimeError("a rank did not deliver its slice of the batch in time (peer-memory result path)")
            return self.host_out[2:2 + W].numpy().copy()
        out = np.empty(W)
        for r in range(self.world):
            out[lo[r]:hi[r]] = self.ret_all[r][:hi[r] - lo[r]].cpu().numpy()
        return out

    def serve(self):
        """Worker loop of the other ranks: take part in every evaluate() of the sampler rank until its shutdown()."""
        assert self.rank != self.src
        n_calls = 0
        while True:
            self._scatter()
            hdr = self.recv[:HDR].cpu().numpy()
            if hdr[0] == CMD_SHUTDOWN:
                return n_calls
            if hdr[0] == CMD_GROW:
                self._setup(int(hdr[1]))
                continue
            self.seq = int(hdr[2])
            self._local(int(hdr[0]), int(hdr[1]))
            n_calls += 1

    def shutdown(self):
        """Sampler rank: release the workers."""
        if self.world > 1 and self.rank == self.src:
            self._command(CMD_SHUTDOWN)
        if self.peer is not None:
            self.torch.cuda.synchronize(self.device)


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
