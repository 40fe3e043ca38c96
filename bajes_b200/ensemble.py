"""
Driver-side helpers for end-to-end runs (BASELINE config 5): a box prior with the attribute surface of
``bajes.inf.prior.Prior`` that ``Posterior`` uses, and a compact affine-invariant ensemble sampler (stretch move,
red/blue split -- the update emcee performs and that bajes' emcee wrapper drives, bajes/inf/sampler/emcee.py:73-185).

They are CALLERS of the hot path, not part of it: like every bajes sampler wrapper, the sampler sees the likelihood
only through ``pool.map(posterior.log_post, points)`` (half an ensemble per call), so with ``BatchedPool`` each
sampler step costs two batched GPU calls.  emcee / dynesty / cpnest / ultranest themselves are not installable
offline; their wrappers reach the GPU through the same ``pool.map`` contract.
"""
from __future__ import annotations

import numpy as np


class BoxPrior(object):
    """Uniform prior on a box + constants (names / bounds / const / in_bounds / log_prior / this_sample /
    prior_transform as in bajes/inf/prior.py:250-324)."""

    def __init__(self, names, bounds, const=None, periodic=None):
        self.names = list(names)
        self.bounds = [list(b) for b in bounds]
        self.const = dict(const or {})
        self.v_names, self.v_funcs, self.v_kwargs = [], [], []
        self.periodics = list(periodic) if periodic is not None else [0] * len(self.names)
        self._lo = np.array([b[0] for b in self.bounds], dtype=float)
        self._hi = np.array([b[1] for b in self.bounds], dtype=float)
        width = self._hi - self._lo
        # a zero-width bound pins the parameter: it carries no volume (and would make the log diverge)
        self._logp = -float(np.sum(np.log(width[width > 0.0])))
        self.ndim = len(self.names)

    def in_bounds(self, x):
        x = np.asarray(x, dtype=float)
        return bool(np.all((x >= self._lo) & (x <= self._hi)))

    def log_prior(self, x):
        return self._logp

    def in_bounds_batch(self, X):
        X = np.asarray(X, dtype=float)
        return np.all((X >= self._lo) & (X <= self._hi), axis=1)

    def log_prior_batch(self, X):
        return np.full(len(X), self._logp)

    def this_sample(self, p):
        return {**p, **self.const}

    def prior_transform(self, u):
        return self._lo + np.asarray(u) * (self._hi - self._lo)

    def prior_transform_batch(self, U):
        return self._lo + np.asarray(U, dtype=float) * (self._hi - self._lo)


class EnsembleSampler(object):
    """Goodman-Weare stretch move on two half-ensembles; every likelihood evaluation goes through
    ``pool.map(posterior.log_post, points)``."""

    def __init__(self, posterior, nwalkers, pool, seed=0, a=2.0):
        if nwalkers % 2:
            raise ValueError("nwalkers must be even")
        self.post = posterior
        self.nwalkers = nwalkers
        self.pool = pool
        self.a = a
        self.rng = np.random.default_rng(seed)
        self.ndim = len(posterior.prior.names)
        self.pos = None
        self.logp = None
        self.n_accept = 0
        self.n_prop = 0
        self.n_eval = 0

    def _eval(self, pts):
        self.n_eval += len(pts)
        return np.asarray(self.pool.map(self.post.log_post, pts), dtype=float)

    def initialize(self, p0):
        self.pos = np.array(p0, dtype=float)
        assert self.pos.shape == (self.nwalkers, self.ndim)
        self.logp = self._eval(self.pos)

    def step(self):
        half = self.nwalkers // 2
        sets = (np.arange(half), np.arange(half, self.nwalkers))
        for s in (0, 1):
            me, other = sets[s], sets[1 - s]
            z = ((self.a - 1.0) * self.rng.random(half) + 1.0) ** 2 / self.a
            partner = self.pos[other[self.rng.integers(0, half, half)]]
            prop = partner + z[:, None] * (self.pos[me] - partner)
            lp = self._eval(prop)
            with np.errstate(invalid="ignore", over="ignore"):
                log_ratio = (self.ndim - 1.0) * np.log(z) + lp - self.logp[me]
            accept = np.log(self.rng.random(half)) < log_ratio
            accept &= np.isfinite(lp)
            self.pos[me[accept]] = prop[accept]
            self.logp[me[accept]] = lp[accept]
            self.n_accept += int(accept.sum())
            self.n_prop += half

    def run(self, nsteps):
        trace = np.empty((nsteps, 3))
        for i in range(nsteps):
            self.step()
            fin = self.logp[np.isfinite(self.logp)]
            trace[i] = (np.max(fin), np.median(fin), np.min(fin)) if len(fin) else (np.nan,) * 3
        return trace

    @property
    def acceptance(self):
        return self.n_accept / max(1, self.n_prop)
