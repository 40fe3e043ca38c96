"""
Pool-shaped adapter: lets bajes' UNCHANGED sampler wrappers reach the batched GPU likelihood.

Every wrapper in bajes/inf/sampler/ evaluates points through ``pool.map(fn, points)`` with
``fn`` one of ``posterior.log_like`` / ``log_post`` / ``log_likeprior`` (SURVEY.md section 3.4:
emcee.py:175, ptmcmc.py:274, ultranest.py:41-42, dynesty.py:140).  ``BatchedPool.map`` recognises those
bound methods, stacks the iterable into one [W, ndim] array, makes ONE batched call and returns the
results in input order -- the contract of ``MPIPool.map`` (bajes/pipe/utils/mpi.py:120-159).  dynesty's
``pool.map(evolve_point, args)`` is recognised as well when ``evolve_point`` is ``LockstepRandomWalk.propose``
(rwalk.py): the chains of one queue are advanced in lock-step.  Anything else is mapped serially.

It subclasses ``multiprocessing.pool.Pool`` only nominally (no worker processes are started) because
ultranest's wrapper insists on ``isinstance(pool, multiprocessing.pool.Pool)`` (ultranest.py:92-96).
"""
from __future__ import annotations

import multiprocessing.pool

import numpy as np

_BATCHED = {
    "log_like": "log_like_batch",
    "log_post": "log_post_batch",
    "log_likeprior": "log_likeprior_batch",
}


class BatchedPool(multiprocessing.pool.Pool):

    def __init__(self, posterior=None, processes=1):   # noqa: D401  (deliberately does NOT call Pool.__init__)
        self._posterior = posterior
        self._processes = int(processes)               # dynesty reads pool._processes (dynesty.py:138-140)
        self._closed = False
        self.n_batched_calls = 0
        self.n_points = 0

    # -- the one method the samplers use ---------------------------------------------------------------
    def map(self, func, iterable, chunksize=None):
        owner = getattr(func, "__self__", None)
        name = getattr(func, "__name__", "")
        if name in ("propose", "sample_rwalk") and hasattr(owner, "propose_batch"):
            # dynesty's evolve_point fan-out (dynesty.py:244, :406-407): all chains advance in lock-step (rwalk.py)
            return owner.propose_batch(list(iterable))
        batched = _BATCHED.get(name)
        if owner is not None and batched is not None and hasattr(owner, batched):
            if isinstance(iterable, np.ndarray) and iterable.ndim == 2:
                X = np.ascontiguousarray(iterable, dtype=np.float64)
            else:
                pts = [np.asarray(p, dtype=np.float64) for p in iterable]
                if len(pts) == 0:
                    return []
                X = np.stack(pts)
            if len(X) == 0:
                return []
            self.n_batched_calls += 1
            self.n_points += len(X)
            res = getattr(owner, batched)(X)
            if name == "log_likeprior":
                lnl, lnp = res
                return list(zip(lnl.tolist(), lnp.tolist()))
            return np.asarray(res).tolist()
        return [func(x) for x in iterable]

    def imap(self, func, iterable, chunksize=1):
        return iter(self.map(func, iterable))

    def starmap(self, func, iterable, chunksize=None):
        return [func(*args) for args in iterable]

    def apply(self, func, args=(), kwds=None):
        return func(*args, **(kwds or {}))

    # -- life-cycle no-ops (bajes/pipe/__init__.py:234-241) -----------------------------------------------
    def close(self):
        self._closed = True

    def terminate(self):
        self._closed = True

    def join(self):
        pass

    def is_master(self):
        return True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        pass

    def __reduce__(self):
        return (BatchedPool, (None, self._processes))
