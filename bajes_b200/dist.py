"""
Multi-GPU evaluation: the walker batch is sharded across the ranks of a ``torch.distributed`` group
(one process per GPU, NCCL over NVLink; gloo on CPU for tests).  Walkers are independent
(``Posterior.log_like`` is a pure function of one point; the reference farms points to workers the same
way, bajes/pipe/utils/mpi.py:120-159), so the data path has NO collective: each walker's whole reduction
happens on one GPU and results are bit-identical for any number of ranks.  The only communication is
the plumbing around it -- and on CUDA that plumbing is peer memory: the parameter block is read by every rank's
prologue kernel, and the W log-likelihoods are stored by every rank's epilogue kernel, straight out of / into the
sampler rank's HBM over NVLink (one scatter + one gather on CPU or without peer access).
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


class _PtrView:
    """A float64 array at a raw device address with the two things a device worker needs from its arguments:
    ``data_ptr()`` and ``shape``.  Used for the sampler rank's buffer as mapped into the OTHER ranks' address space
    (torch would attribute that memory to the sampler rank's device)."""

    def __init__(self, ptr: int, n: int, shape=None):
        self.ptr, self.n = int(ptr), int(n)
        self.shape = tuple(shape) if shape is not None else (self.n,)

    def data_ptr(self) -> int:
        return self.ptr

    def __len__(self):
        return self.shape[0]


HDR = 3                     # per-rank header of the scattered block: rows of this rank, offset of its first row, sequence
CMD_RUN, CMD_SHUTDOWN, CMD_GROW = 0.0, -1.0, -2.0
X_FLAG0 = 32                # flag slots 32 + r: "rank r's rows of X are in the sampler rank's HBM" (slots r: "rank r is done")


class ShardedEvaluator:
    """``evaluate(X)``: rank ``src`` (the sampler) supplies X [W, ndim] and gets logL [W] back; every rank evaluates a
    contiguous slice.  Two transports:

    **peer memory** (CUDA, the default when every GPU can map the sampler rank's memory): NO collective per call.
    The sampler rank owns one device buffer [flags | logL | X]; the other ranks map it through a CUDA IPC handle.
      out    the sampler rank writes the per-rank (rows, first row) table to a small host shared-memory block and
             enqueues ONE host-to-device copy of X into its buffer followed by a flag store; the other ranks' host loops
             see the new sequence number, enqueue a one-block kernel that waits for that flag (a remote read over
             NVLink) and then their likelihood kernels, whose prologue reads its rows of X straight out of the sampler
             rank's HBM;
      back   every rank's epilogue kernel stores its logL straight into the sampler rank's HBM, then publishes a
             system-scope flag; a one-block kernel on the sampler rank's stream waits for all flags, the device-to-
             host copy follows in stream order.  One host synchronisation on the sampler rank, none elsewhere.
    Walkers are independent, so this IS the whole exchange of the path: compute and "collective" are the same kernels.

    **collective** (CPU / gloo tests, or GPUs without peer access): one ``scatter`` of fixed-capacity per-rank blocks
    [rows, first row, sequence | rows of X] and one ``gather`` of the results.

    ``local_eval(X_slice) -> logL_slice`` is the per-rank host worker (CPU tests pass any pure function);
    ``local_eval_device(x_dev [w, ndim], out_dev [w])`` evaluates a device-resident slice in place on the current
    stream; both arguments only need ``data_ptr()`` and ``shape`` -- they may be views of the sampler rank's memory.
    The capacity grows on demand (collectively).  The other ranks sit in ``serve()`` (the role of MPIPool.wait,
    bajes/pipe/utils/mpi.py:91-117)."""

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
        self.shm = None
        self.want_peer = (self.device.type == "cuda" and local_eval_device is not None) if peer is None else bool(peer)
        self.capacity = 0
        if self.world > 1:
            assert self.world <= X_FLAG0
            self._setup(int(capacity))

    # -- collective set-up (every rank, same capacity) ------------------------------------------------------------
    def _release(self):
        if self.peer is not None:
            self.torch.cuda.synchronize(self.device)
            # the importers unmap before the owner frees
            if self.rank == self.src:
                self.dist.barrier(group=self.group)
            self.peer.close()
            if self.rank != self.src:
                self.dist.barrier(group=self.group)
            self.peer = None
        if self.shm is not None:
            self.hdr = None
            self.shm.close()
            if self.rank == self.src:
                self.shm.unlink()
            self.shm = None

    def _setup(self, capacity: int):
        torch, dist = self.torch, self.dist
        self._release()
        self.capacity = int(capacity)
        cuda = self.device.type == "cuda"
        if self.want_peer:
            from multiprocessing import shared_memory
            from . import engine
            ok = torch.ones(1, dtype=torch.int32, device=self.device)
            info = [None]
            try:
                if self.rank == self.src:
                    self.peer = engine.PeerBuffer.create(self.device.index, self.capacity * (1 + self.ndim))
                    self.shm = shared_memory.SharedMemory(create=True, size=8 * 4 * self.world)
                    info[0] = (self.peer.handle, self.shm.name)
            except Exception:
                ok.zero_()
            dist.broadcast_object_list(info, src=self.src, group=self.group)
            if self.rank != self.src:
                try:
                    self.peer = engine.PeerBuffer.open(self.device.index, info[0][0], self.capacity * (1 + self.ndim))
                    self.shm = shared_memory.SharedMemory(name=info[0][1])
                    try:                                 # the owner unlinks it; keep this process's tracker out of it
                        from multiprocessing import resource_tracker
                        resource_tracker.unregister(self.shm._name, "shared_memory")
                    except Exception:
                        pass
                except Exception:
                    ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:                      # some pair of GPUs has no peer access: collectives instead
                self._release()
            else:
                from .engine import PEER_FLAGS
                # host header, one row per rank: sequence, rows, first row, command
                self.hdr = np.ndarray((self.world, 4), dtype=np.float64, buffer=self.shm.buf)
                if self.rank == self.src:
                    self.hdr[:] = 0.0
                    self.out_all = torch.as_tensor(_RawCuda(self.peer.base + PEER_FLAGS * 8, 2 + self.capacity),
                                                   device=self.device)         # [status, pad, logL...]
                    self.x_all = torch.as_tensor(_RawCuda(self.peer.data_ptr + 8 * self.capacity,
                                                          self.capacity * self.ndim), device=self.device)
                    self.x_pin = torch.empty(self.capacity * self.ndim, dtype=torch.float64, pin_memory=True)
                    self.host_out = torch.empty(2 + self.capacity, dtype=torch.float64, pin_memory=True)
                self.seq = 0
                dist.barrier(group=self.group)           # the header is zeroed before anyone polls it
                return
        # collective transport
        self.cap_rank = -(-self.capacity // self.world)
        self.slot = HDR + self.cap_rank * self.ndim
        self.recv = torch.empty(self.slot, dtype=torch.float64, device=self.device)
        if self.rank == self.src:
            self.stage = torch.empty((self.world, self.slot), dtype=torch.float64, pin_memory=cuda)
            self.send = self.stage if not cuda else torch.empty((self.world, self.slot), dtype=torch.float64,
                                                                 device=self.device)
        self.ret = torch.empty(self.cap_rank, dtype=torch.float64, device=self.device)
        self.ret_all = ([torch.empty(self.cap_rank, dtype=torch.float64, device=self.device)
                         for _ in range(self.world)] if self.rank == self.src else None)

    # -- peer-memory transport ----------------------------------------------------------------------------------------
    def _peer_local(self, n: int, lo: int):
        """this rank's rows [lo, lo + n): X and logL both live in the sampler rank's buffer"""
        if n > 0:
            x = _PtrView(self.peer.data_ptr + 8 * (self.capacity + lo * self.ndim), n * self.ndim, (n, self.ndim))
            self.local_eval_device(x, _PtrView(self.peer.data_ptr + 8 * lo, n, (n,)))

    def _peer_evaluate(self, X) -> np.ndarray:
        torch = self.torch
        W = len(X)
        self.seq += 1
        lo, hi = shard_bounds(W, self.world)
        stream = torch.cuda.current_stream(self.device)
        # the header first: the other ranks enqueue their kernels behind a wait on THEIR slice's flag while it travels
        for r in range(self.world):
            self.hdr[r, 1], self.hdr[r, 2], self.hdr[r, 3] = hi[r] - lo[r], lo[r], CMD_RUN
        self.hdr[:, 0] = self.seq                        # published last
        # X into the sampler rank's HBM, slice by slice (the other ranks' first, each followed by its flag), staged through
        # pinned memory in pieces so that the host copy of one piece overlaps the DMA of the previous one
        pinned = isinstance(X, torch.Tensor)              # evaluate() only lets pinned float64 tensors through as such
        src_ptr = X.data_ptr() if pinned else X.ctypes.data
        order = [q for q in range(self.world) if q != self.src] + [self.src]
        from .engine import PEER_HEADER_BYTES
        self.peer.scatter(PEER_HEADER_BYTES + 8 * self.capacity, src_ptr, 0 if pinned else self.x_pin.data_ptr(),
                          [8 * lo[r] * self.ndim for r in order], [8 * hi[r] * self.ndim for r in order],
                          [X_FLAG0 + r for r in order], self.seq, stream.cuda_stream)
        self._peer_local(hi[self.src] - lo[self.src], lo[self.src])
        others = [r for r in range(self.world) if r != self.src]
        first, last = min(others), max(others)
        if first < self.src < last:
            self.peer.wait(first, self.src - first, self.seq, stream=stream.cuda_stream)
            self.peer.wait(self.src + 1, last - self.src, self.seq, stream=stream.cuda_stream)
        else:
            self.peer.wait(first, last - first + 1, self.seq, stream=stream.cuda_stream)
        self.host_out[:2 + W].copy_(self.out_all[:2 + W], non_blocking=True)
        stream.synchronize()
        if int(self.host_out[:1].view(torch.int64)[0]) != 0:
            raise RuntimeError("a rank did not deliver its slice of the batch in time (peer-memory result path)")
        return self.host_out[2:2 + W].numpy().copy()

    def _peer_command(self, cmd: float, arg: float = 0.0):
        self.seq += 1
        self.hdr[:, 1], self.hdr[:, 2], self.hdr[:, 3] = arg, 0.0, cmd
        self.hdr[:, 0] = self.seq

    def _peer_serve(self):
        import time
        torch = self.torch
        stream = torch.cuda.current_stream(self.device)
        n_calls, idle = 0, 0
        while True:
            seq = self.hdr[self.rank, 0]
            if seq == self.seq:
                idle += 1
                if idle > 20000:                         # a sampler thinking on the host: stop burning a core
                    time.sleep(0.0002)
                continue
            idle = 0
            self.seq = int(seq)
            n, lo, cmd = int(self.hdr[self.rank, 1]), int(self.hdr[self.rank, 2]), self.hdr[self.rank, 3]
            if cmd == CMD_SHUTDOWN:
                return n_calls, None
            if cmd == CMD_GROW:
                return n_calls, n
            # the parameter block is in the sampler rank's HBM once its flag carries this sequence number
            self.peer.wait(X_FLAG0 + self.rank, 1, self.seq, stream=stream.cuda_stream)
            self._peer_local(n, lo)
            self.peer.signal(self.rank, self.seq, stream.cuda_stream)
            n_calls += 1

    # -- collective transport -----------------------------------------------------------------------------------------
    def _scatter(self):
        self.dist.scatter(self.recv, scatter_list=list(self.send.unbind(0)) if self.rank == self.src else None,
                          src=self.src, group=self.group)

    def _command(self, cmd: float, arg: float = 0.0):
        """sampler rank: a header-only block for every rank (shutdown / grow)"""
        if self.peer is not None:
            return self._peer_command(cmd, arg)
        self.stage[:, 0] = cmd
        self.stage[:, 1] = arg
        if self.send is not self.stage:
            self.send[:, :HDR].copy_(self.stage[:, :HDR], non_blocking=True)
        self._scatter()

    def _local(self, n: int):
        """evaluate the n rows that sit in self.recv and hand them to the gather"""
        torch = self.torch
        rows = self.recv[HDR:HDR + n * self.ndim].view(n, self.ndim)
        if n > 0:
            if self.local_eval_device is not None:
                self.local_eval_device(rows, self.ret[:n])
            else:
                out = np.asarray(self.local_eval(rows.cpu().numpy()), dtype=np.float64)
                self.ret[:n] = torch.as_tensor(out).to(self.device)
        self.dist.gather(self.ret, gather_list=self.ret_all, dst=self.src, group=self.group)

    def evaluate(self, X) -> np.ndarray:
        """X: float64 [W, ndim], a NumPy array or (peer transport) a pinned torch tensor, which skips the staging copy"""
        if not (self.peer is not None and isinstance(X, self.torch.Tensor) and X.is_pinned()
                and X.dtype == self.torch.float64 and X.is_contiguous()):
            X = np.ascontiguousarray(X.numpy() if isinstance(X, self.torch.Tensor) else X, dtype=np.float64)
        if self.world == 1:
            return np.asarray(self.local_eval(X))
        assert self.rank == self.src, "evaluate() is the sampler rank's call; the other ranks run serve()"
        W = len(X)
        if W > self.capacity:
            new_cap = max(W, 2 * self.capacity)
            self._command(CMD_GROW, float(new_cap))
            self._setup(new_cap)
        if self.peer is not None:
            return self._peer_evaluate(X)
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
        self._local(hi[self.src] - lo[self.src])
        out = np.empty(W)
        for r in range(self.world):
            out[lo[r]:hi[r]] = self.ret_all[r][:hi[r] - lo[r]].cpu().numpy()
        return out

    def serve(self):
        """Worker loop of the other ranks: take part in every evaluate() of the sampler rank until its shutdown()."""
        assert self.rank != self.src
        n_calls = 0
        while True:
            if self.peer is not None:
                k, grow = self._peer_serve()
                n_calls += k
                if grow is None:
                    self._release()
                    return n_calls
                self._setup(int(grow))
                continue
            self._scatter()
            hdr = self.recv[:HDR].cpu().numpy()
            if hdr[0] == CMD_SHUTDOWN:
                return n_calls
            if hdr[0] == CMD_GROW:
                self._setup(int(hdr[1]))
                continue
            self.seq = int(hdr[2])
            self._local(int(hdr[0]))
            n_calls += 1

    def shutdown(self):
        """Sampler rank: release the workers."""
        if self.world > 1 and self.rank == self.src:
            self._command(CMD_SHUTDOWN)
            self._release()

    def __del__(self):
        try:
            if self.shm is not None:
                self.hdr = None
                self.shm.close()
                if self.rank == self.src:
                    self.shm.unlink()
                self.shm = None
        except Exception:
            pass


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
