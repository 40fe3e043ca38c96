"""
Lock-step batched random walk for dynesty.

bajes replaces dynesty's ``evolve_point`` with ``BajesDynestyProposal.propose`` (bajes/inf/sampler/dynesty.py:244,
:406-407) and dynesty fans ``queue_size`` such calls out through ``pool.map`` (dynesty.py:212-216): each call walks ONE
chain of up to ``maxmcmc`` proposals, one likelihood evaluation at a time (dynesty.py:503-614).  A GPU wants the
opposite nesting: all chains advanced together, one proposal per chain per batched likelihood call.

``LockstepRandomWalk.propose_batch(args_list)`` does exactly that and returns, per chain, the tuple
``(u, v, logl, ncall, blob)`` of ``sample_rwalk`` (dynesty.py:612-614).  Every rule of the per-chain state machine is
the reference's:

* proposal: direction uniform on the n-sphere, radius ``rand()**(1/n)``, mapped by ``axes``, scaled by ``scale``
  (dynesty.py:537-549); periodic wrap, reflection (:45-49, :551-556), unit-cube check (:27-42, :559);
* a proposal that leaves the cube costs no likelihood call, bumps ``nfail`` and -- once the chain has accepted at
  least one point -- repeats the last chain entry (:559-568);
* accept iff ``logl_prop > loglstar`` (:574); the chain only starts to be recorded at its first accept (:581-586);
* after ``walks`` evaluated steps and 2 accepts the autocorrelation estimate is refreshed with
  ``estimate_nmcmc(accept / (accept + reject + nfail), walks, maxmcmc)`` (:589-590, inf/utils.py:140-161), and the
  loop runs while ``0.5 * nact * act >= len(chain)`` (:533) or until ``accept + reject > maxmcmc`` (:593-595);
* the returned point is drawn uniformly from the chain beyond ``int(0.5 * nact * act)`` (:598-602); a chain that
  never got there returns a fresh prior draw (:603-607).

Chains are independent, so advancing them in lock-step changes nothing but the order in which random numbers are
consumed.  With ``draws="chain"`` each chain owns a ``numpy.random.RandomState``; a chain then reproduces, bit for bit,
what the reference's ``sample_rwalk`` returns after ``numpy.random.seed`` with the same seed (this is how the golden
test pins the state machine).  ``draws="vector"`` (default) draws for all active chains at once from one generator.

``BatchedPool.map(walk.propose, args_list)`` routes here (pool.py), so a dynesty sampler whose ``evolve_point`` is
``walk.propose`` and whose pool is a ``BatchedPool`` runs lock-step without further changes.
"""
from __future__ import annotations

import numpy as np

__all__ = ["LockstepRandomWalk", "estimate_nmcmc", "reflect", "unitcheck"]


def estimate_nmcmc(accept_ratio, old_act, maxmcmc, safety=5, tau=None):
    """bajes/inf/utils.py:140-161, element-wise over arrays of acceptance ratios."""
    accept_ratio = np.asarray(accept_ratio, dtype=np.float64)
    if tau is None:
        tau = maxmcmc / safety
    with np.errstate(divide="ignore"):
        moving = np.minimum((1.0 - 1.0 / tau) * old_act + (safety / tau) * (2.0 / accept_ratio - 1.0), float(maxmcmc))
    exact = np.where(accept_ratio == 0.0, (1.0 + 1.0 / tau) * old_act, moving)
    return np.maximum(safety, exact.astype(np.int64))


def reflect(u):
    """dynesty.py:45-49 on the last axis (out of place)."""
    u = np.asarray(u, dtype=np.float64)
    even = np.mod(u, 2) < 1
    return np.where(even, np.mod(u, 1), 1 - np.mod(u, 1))


def unitcheck(U, nonbounded=None):
    """dynesty.py:27-42, row-wise.  ``nonbounded`` is dynesty's boolean mask (True = hard [0, 1] edges)."""
    U = np.atleast_2d(U)
    if nonbounded is not None and not np.any(nonbounded):
        nonbounded = None
    if nonbounded is None:
        return (U.min(axis=1) > 0) & (U.max(axis=1) < 1)
    nb = np.asarray(nonbounded, dtype=bool)
    ok = np.ones(len(U), dtype=bool)
    if nb.any():
        ok &= (U[:, nb].min(axis=1) > 0) & (U[:, nb].max(axis=1) < 1)
    if (~nb).any():
        ok &= (U[:, ~nb].min(axis=1) > -0.5) & (U[:, ~nb].max(axis=1) < 1.5)
    return ok


class _Chains(object):
    """Recorded part of all chains: per chain the accepted points and the chain length at which each was appended
    (between two accepts the chain repeats its last entry, so that is all there is to store).  Array-backed: appending
    the accepts of one proposal round is a handful of fancy-indexed assignments, whatever the number of chains."""

    def __init__(self, m, n, cap=32):
        self.m, self.n, self.cap = m, n, cap
        self.u = np.empty((m, cap, n))
        self.v = np.empty((m, cap, n))
        self.l = np.empty((m, cap))
        self.pos = np.zeros((m, cap), dtype=np.int64)
        self.count = np.zeros(m, dtype=np.int64)

    def _grow(self):
        for name in ("u", "v", "l", "pos"):
            old = getattr(self, name)
            new = np.zeros((self.m, 2 * self.cap) + old.shape[2:], dtype=old.dtype)
            new[:, :self.cap] = old
            setattr(self, name, new)
        self.cap *= 2

    def append(self, idx, u, v, l, pos):
        if len(idx) == 0:
            return
        k = self.count[idx]
        if k.max() >= self.cap:
            self._grow()
        self.u[idx, k], self.v[idx, k], self.l[idx, k], self.pos[idx, k] = u, v, l, pos
        self.count[idx] = k + 1

    def at(self, c, step):
        """Entry ``step`` of chain ``c``: the last accept appended at or before it."""
        j = int(np.searchsorted(self.pos[c, :self.count[c]], step, side="right")) - 1
        return self.u[c, j], self.v[c, j], self.l[c, j]


class LockstepRandomWalk(object):
    """Batched drop-in for ``BajesDynestyProposal`` (dynesty.py:487-501) restricted to ``sample_rwalk``.

    posterior : object with ``log_like_batch(V) -> [m]`` and ``prior_transform`` (per point; a
                ``prior_transform_batch`` is used when present) -- e.g. ``bajes_b200.Posterior``.
    """

    def __init__(self, maxmcmc, walks=100, nact=5.0, posterior=None, loglike_batch=None, prior_transform_batch=None,
                 draws="vector", seed=None):
        self.maxmcmc, self.walks, self.nact = maxmcmc, walks, nact
        self.use_slice = False
        if posterior is not None:
            loglike_batch = loglike_batch or posterior.log_like_batch
            if prior_transform_batch is None:
                prior_transform_batch = getattr(posterior, "prior_transform_batch", None)
                if prior_transform_batch is None:
                    pt = posterior.prior_transform
                    prior_transform_batch = lambda U: np.stack([np.asarray(pt(u), dtype=np.float64) for u in U])  # noqa: E731
        if loglike_batch is None or prior_transform_batch is None:
            raise ValueError("LockstepRandomWalk needs a batched likelihood and a prior transform")
        self._loglike, self._ptform = loglike_batch, prior_transform_batch
        if draws not in ("vector", "chain"):
            raise ValueError("draws must be 'vector' or 'chain'")
        self.draws = draws
        self._rng = np.random.default_rng(seed)
        self.n_batched_calls = 0
        self.n_evaluations = 0

    # -- the per-point entry keeps working (a batch of one) ------------------------------------------------
    def propose(self, args):
        return self.propose_batch([args])[0]

    sample_rwalk = propose

    # -- helpers ---------------------------------------------------------------------------------------------
    def _evaluate(self, U):
        V = np.asarray(self._ptform(U), dtype=np.float64)
        L = np.asarray(self._loglike(np.ascontiguousarray(V)), dtype=np.float64)
        self.n_batched_calls += 1
        self.n_evaluations += len(U)
        return V, L

    @staticmethod
    def _unzip(args):
        if len(args) == 7:                                  # dynesty < 1.2 (dynesty.py:512-513)
            u, loglstar, axes, scale, _pt, _ll, kwargs = args
            seedseq = None
        else:
            u, loglstar, axes, scale, _pt, _ll, seedseq, kwargs = args
        return np.asarray(u, dtype=np.float64), float(loglstar), np.asarray(axes, dtype=np.float64), float(scale), \
            seedseq, (kwargs or {})

    def propose_batch(self, args_list, seeds=None):
        """Advance all chains of ``args_list`` (the tuples dynesty hands to ``evolve_point``) in lock-step."""
        m = len(args_list)
        if m == 0:
            return []
        un = [self._unzip(a) for a in args_list]
        n = len(un[0][0])
        u = np.stack([x[0] for x in un])
        loglstar = np.array([x[1] for x in un])
        axes = np.stack([x[2] for x in un])
        scale = np.array([x[3] for x in un])
        kw = un[0][5]
        nonbounded, periodic, reflective = kw.get("nonbounded"), kw.get("periodic"), kw.get("reflective")
        if self.draws == "chain":
            if seeds is None:
                seeds = [x[4] for x in un]
            states = [s if isinstance(s, np.random.RandomState) else np.random.RandomState(
                s.generate_state(4) if hasattr(s, "generate_state") else s) for s in seeds]

        v = np.full((m, n), np.nan)
        logl = np.full(m, np.nan)
        accept = np.zeros(m, dtype=np.int64)
        reject = np.zeros(m, dtype=np.int64)
        nfail = np.zeros(m, dtype=np.int64)
        nlist = np.zeros(m, dtype=np.int64)
        act = np.full(m, np.inf)
        active = np.ones(m, dtype=bool)
        chains = _Chains(m, n)
        old_act = float(self.walks)

        while True:
            active &= 0.5 * self.nact * act >= nlist                     # dynesty.py:533
            idx = np.flatnonzero(active)
            if len(idx) == 0:
                break
            k = len(idx)
            if self.draws == "chain":
                # per chain, with the reference's own operations (1-D norm, matrix-vector dot) so that a chain is
                # bit-identical to sample_rwalk fed the same stream
                du = np.empty((k, n))
                for j, c in enumerate(idx):
                    drhat = states[c].randn(n)
                    drhat /= np.linalg.norm(drhat)
                    du[j] = np.dot(axes[c], drhat * states[c].rand() ** (1.0 / n))
            else:
                drhat = self._rng.standard_normal((k, n))
                drhat /= np.linalg.norm(drhat, axis=1)[:, None]
                dr = drhat * (self._rng.random(k) ** (1.0 / n))[:, None]
                du = np.einsum("kij,kj->ki", axes[idx], dr)
            u_prop = u[idx] + scale[idx, None] * du
            if periodic is not None:
                u_prop[:, periodic] = np.mod(u_prop[:, periodic], 1)
            if reflective is not None:
                u_prop[:, reflective] = reflect(u_prop[:, reflective])
            ok = unitcheck(u_prop, nonbounded)

            bad = idx[~ok]                                               # dynesty.py:561-568
            nfail[bad] += 1
            nlist[bad] += accept[bad] > 0

            good = idx[ok]
            if len(good):
                v_prop, l_prop = self._evaluate(u_prop[ok])
                acc = l_prop > loglstar[good]                            # dynesty.py:574
                a_idx = good[acc]
                u[a_idx], v[a_idx], logl[a_idx] = u_prop[ok][acc], v_prop[acc], l_prop[acc]
                accept[a_idx] += 1
                chains.append(a_idx, u[a_idx], v[a_idx], logl[a_idx], nlist[a_idx])
                nlist[a_idx] += 1
                r_idx = good[~acc]
                reject[r_idx] += 1
                nlist[r_idx] += accept[r_idx] > 0
                tried = accept[good] + reject[good]
                upd = good[(tried > self.walks) & (accept[good] > 1)]   # dynesty.py:589-590
                if len(upd):
                    ratio = accept[upd] / (accept[upd] + reject[upd] + nfail[upd])
                    act[upd] = estimate_nmcmc(ratio, old_act, self.maxmcmc)
                active[good[tried > self.maxmcmc]] = False               # dynesty.py:593-595

        # -- pick the returned point (dynesty.py:598-607) ---------------------------------------------------------
        out_u, out_v, out_l = u.copy(), v.copy(), logl.copy()
        fresh = []
        for c in range(m):
            lo = int(0.5 * self.nact * act[c]) if np.isfinite(act[c]) else None
            if lo is not None and lo < nlist[c]:
                pick = states[c].randint(lo, nlist[c]) if self.draws == "chain" else int(self._rng.integers(lo, nlist[c]))
                out_u[c], out_v[c], out_l[c] = chains.at(c, pick)
            else:
                fresh.append(c)
        if fresh:
            if self.draws == "chain":
                uf = np.stack([states[c].uniform(size=n) for c in fresh])
            else:
                uf = self._rng.uniform(size=(len(fresh), n))
            vf, lf = self._evaluate(uf)
            out_u[fresh], out_v[fresh], out_l[fresh] = uf, vf, lf
        res = []
        for c in range(m):
            blob = {"accept": int(accept[c]), "reject": int(reject[c]), "fail": int(nfail[c]), "scale": float(scale[c])}
            res.append((out_u[c], out_v[c], float(out_l[c]), int(accept[c] + reject[c]), blob))
        return res
