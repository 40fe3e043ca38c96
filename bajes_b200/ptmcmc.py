"""
Parallel-tempered ensemble sampler, batch-native: a restatement of the ensemble kernel of bajes' own PTMCMC
(``_PTMCMC``, bajes/inf/sampler/ptmcmc.py:184-470; temperature ladder :37-146; stretch move
bajes/inf/sampler/proposal.py:92-117; boundary handling bajes/inf/utils.py:165-198) that evaluates every half-ensemble of
every temperature with ONE batched call ``posterior.log_likeprior_batch(points)`` instead of
``pool.map(posterior.log_likeprior, points)`` (:272-281).

It is a CALLER of the hot path, written so that config C5 -- "the C2 data driven by a sampler"
-- can run where the reference tree cannot travel (the GPU box).  Two random-number modes:

  ``rng="reference"``  draws exactly what the reference draws, in its order and from its streams -- the proposal choice, the
                       acceptance uniforms and the temperature swaps from ``random`` (a ``numpy.random.RandomState``), the
                       stretch move from the GLOBAL legacy stream, one ``randint`` and one ``rand`` per walker
                       (proposal.py:96-97; a quirk of the reference: its moves ignore the sampler's generator).  Chains are
                       bit-identical to the unmodified ``_PTMCMC`` with ``props={'str': 1}``
                       (tests/test_ptmcmc.py: against the reference itself where its tree exists, and against
                       tests/golden/ptmcmc_trace.npz, recorded from it, everywhere).
  ``rng="vector"``     the same distributions drawn in whole arrays from ``random`` (a Generator or RandomState): the mode
                       for large ensembles, where 10^4 Python-level draws per step would cost more than the likelihood.

Kept from the reference: walkers of one temperature are updated in two halves against the complementary half; proposals
are moved into the prior bounds (periodic wrap / reflection) BEFORE evaluation; ``(lnL, lnp) = (0, -inf)`` outside the
bounds; tempered posterior ``beta * lnL + lnp`` with NaN -> -inf; ``(ndim - 1) ln z`` Jacobian of the stretch move; swaps
between adjacent temperatures over two independent random permutations, from the hottest pair down; the geometric
ladder of ``default_beta_ladder``.  Not restated: the other moves of ``BajesPTMCMCProposal`` (eigen-vector, differential
evolution, KDE, walk, prior, slice, GW-targeted) -- the reference's own trace fixture uses the stretch move alone.
"""
from __future__ import annotations

import numpy as np

# 25 %-swap-acceptance temperature steps per dimension (ptmcmc.py:84-110)
_TSTEP = np.array([25.2741, 7., 4.47502, 3.5236, 3.0232, 2.71225, 2.49879, 2.34226, 2.22198, 2.12628, 2.04807, 1.98276,
                   1.92728, 1.87946, 1.83774, 1.80096, 1.76826, 1.73895, 1.7125, 1.68849, 1.66657, 1.64647, 1.62795, 1.61083,
                   1.59494, 1.58014, 1.56632, 1.55338, 1.54123, 1.5298, 1.51901, 1.50881, 1.49916, 1.49, 1.4813, 1.47302,
                   1.46512, 1.45759, 1.45039, 1.4435, 1.4369, 1.43056, 1.42448, 1.41864, 1.41302, 1.40761, 1.40239, 1.39736,
                   1.3925, 1.38781, 1.38327, 1.37888, 1.37463, 1.37051, 1.36652, 1.36265, 1.35889, 1.35524, 1.3517, 1.34825,
                   1.3449, 1.34164, 1.33847, 1.33538, 1.33236, 1.32943, 1.32656, 1.32377, 1.32104, 1.31838, 1.31578, 1.31325,
                   1.31076, 1.30834, 1.30596, 1.30364, 1.30137, 1.29915, 1.29697, 1.29484, 1.29275, 1.29071, 1.2887, 1.28673,
                   1.2848, 1.28291, 1.28106, 1.27923, 1.27745, 1.27569, 1.27397, 1.27227, 1.27061, 1.26898, 1.26737, 1.26579,
                   1.26424, 1.26271, 1.26121, 1.25973])


def default_beta_ladder(ndim, ntemps=None, Tmax=None):
    """Geometric ladder of inverse temperatures (ptmcmc.py:37-146), same argument rules and error messages."""
    if type(ndim) != int or ndim < 1:
        raise ValueError('Invalid number of dimensions specified.')
    if ntemps is None and Tmax is None:
        raise ValueError('Must specify one of ``ntemps`` and ``Tmax``.')
    if Tmax is not None and Tmax <= 1:
        raise ValueError('``Tmax`` must be greater than 1.')
    if ntemps is not None and (type(ntemps) != int or ntemps < 1):
        raise ValueError('Invalid number of temperatures specified.')
    if ndim > _TSTEP.shape[0]:
        tstep = 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim)
    else:
        tstep = _TSTEP[ndim - 1]
    append_inf = False
    if Tmax == np.inf:
        append_inf = True
        Tmax = None
        ntemps = ntemps - 1
    if ntemps is not None:
        if Tmax is None:
            Tmax = tstep ** (ntemps - 1)
    else:
        if Tmax is None:
            raise ValueError('Must specify at least one of ``ntemps'' and finite ``Tmax``.')
        ntemps = int(np.log(Tmax) / np.log(tstep) + 2)
    betas = np.logspace(0, -np.log10(Tmax), ntemps)
    if append_inf:
        betas = np.concatenate((betas, [0]))
    return betas


def apply_bounds_batch(Q, periodic, bounds):
    """Row-wise ``apply_bounds`` (utils.py:168-198): periodic wrap or reflection of every out-of-bounds coordinate, element
    by element with the reference's own expressions (so the bits agree); in-bounds coordinates are untouched."""
    Q = np.array(Q, dtype=np.float64, copy=True)
    b = np.asarray(bounds, dtype=np.float64)
    if not (np.any(Q < b[:, 0]) or np.any(Q > b[:, 1])):         # the common case: nothing to move (list_in_bounds, :171)
        return Q
    for i, (per, (lo, up)) in enumerate(zip(periodic, bounds)):
        x = Q[..., i]
        below, above = x < lo, x > up
        if per:
            x[below] = up - np.abs(x[below] - lo) % (up - lo)
            x[above] = (x[above] - lo) % (up - lo) + lo
        else:
            x[below] = 2. * lo - x[below]
            x[above] = 2. * up - x[above]
    return Q


class PTEnsembleSampler(object):
    """``sample(iterations)`` advances ``ntemps x nwalkers`` walkers; histories as in the reference
    (``_chain_history [ntemps, nwalkers, iterations, dim]``, ``_loglikelihood_history``, ``_logposterior_history``)."""

    def __init__(self, nwalkers, posterior, ntemps=None, Tmax=None, betas=None, stretch=2.0, random=None, rng="reference",
                 adaptation_lag=10000, adaptation_time=100, refocus_every=0, focus_fn=None):
        """``refocus_every`` / ``focus_fn`` (no counterpart in the reference): every that many iterations the best point of
        the cold chain is handed to ``focus_fn(point_row)`` -- by default ``posterior.like.focus(row, names, const)`` when the
        likelihood has one (bajes_b200.GWLikelihood.focus: the focus table of DESIGN.md section 4.9).  It moves no random
        number and changes no result beyond the likelihood's parity tolerance; 0 = never."""
        if nwalkers % 2 != 0:
            raise ValueError('The number of walkers must be even.')
        if rng not in ("reference", "vector"):
            raise ValueError("rng must be 'reference' or 'vector'")
        self.nwalkers = nwalkers
        self.post = posterior
        self.dim = len(posterior.prior.names)
        self.a = float(stretch)
        self.rng_mode = rng
        self._random = np.random if random is None else random
        self.adaptation_lag, self.adaptation_time = adaptation_lag, adaptation_time
        self._period_reflect_list = list(posterior.prior.periodics)
        self._bounds = [list(b) for b in posterior.prior.bounds]
        if betas is not None:
            self._betas = np.array(betas).copy()
        else:
            self._betas = default_beta_ladder(self.dim, ntemps=ntemps, Tmax=Tmax)
        self._betas[::-1].sort()                      # ascending in temperature (ptmcmc.py:212)
        self.ntemps = len(self._betas)
        self._chain_history, self._loglikelihood_history, self._logposterior_history = [], [], []
        self._acceptance_history, self._tswap_acceptance_history = [], []
        self._time = 0
        self._p0 = self._logposterior0 = self._loglikelihood0 = None
        self.nswap = np.zeros(self.ntemps)
        self.nswap_accepted = np.zeros(self.ntemps)
        self.nprop = np.zeros((self.ntemps, self.nwalkers))
        self.nprop_accepted = np.zeros((self.ntemps, self.nwalkers))
        self.n_batched_calls = 0
        self.n_points = 0
        self.refocus_every = int(refocus_every)
        self.n_refocus = 0
        if focus_fn is None and self.refocus_every > 0 and hasattr(getattr(posterior, "like", None), "focus"):
            prior = posterior.prior
            focus_fn = lambda row: posterior.like.focus(row, names=list(prior.names), const=dict(getattr(prior, "const", {})))
        self._focus_fn = focus_fn

    # ---- likelihood: ONE batched call per half-ensemble of all temperatures (ptmcmc.py:272-281) -----------------------
    def _evaluate(self, ps):
        flat = ps.reshape((-1, self.dim))
        lnl, lnp = self.post.log_likeprior_batch(flat)
        self.n_batched_calls += 1
        self.n_points += len(flat)
        return (np.asarray(lnl, dtype=float).reshape((self.ntemps, -1)), np.asarray(lnp, dtype=float).reshape((self.ntemps, -1)))

    def _tempered_likelihood(self, logl, betas=None):
        betas = (self._betas if betas is None else betas).reshape((-1, 1))
        with np.errstate(invalid='ignore'):
            loglT = logl * betas
        loglT[np.isnan(loglT)] = -np.inf
        return loglT

    # ---- the stretch move (proposal.py:92-117) ---------------------------------------------------------------------------
    def _stretch(self, s, c):
        """s [ntemps, n, dim] walkers to update, c [ntemps, n, dim] the complementary half -> (q, (dim - 1) ln z)"""
        nt, n, _ = s.shape
        nc = c.shape[1]
        if self.rng_mode == "reference":
            # the reference first draws which proposal every temperature uses (a choice over its one-entry cycle still
            # consumes ntemps uniforms of the sampler's generator: ptmcmc.py:179), then, per temperature and walker, one
            # randint and one rand from the GLOBAL stream
            self._random.choice(1, size=nt, p=np.array([1.0]))
            ri = np.empty((nt, n), dtype=np.int64)
            u = np.empty((nt, n))
            for t in range(nt):
                for k in range(n):
                    ri[t, k] = np.random.randint(nc)
                    u[t, k] = np.random.rand()
        else:
            if hasattr(self._random, "integers"):
                ri = self._random.integers(0, nc, size=(nt, n))
            else:
                ri = self._random.randint(nc, size=(nt, n))
            u = self._random.random_sample((nt, n)) if hasattr(self._random, "random_sample") else self._random.random((nt, n))
        zz = ((self.a - 1.0) * u + 1) ** 2.0 / self.a
        partner = np.take_along_axis(c, ri[:, :, None], axis=1)
        q = partner - (partner - s) * zz[:, :, None]
        return q, (self.dim - 1.0) * np.log(zz)

    def _propose(self, p, logpost, logl):
        for j in (0, 1):
            jupdate, jsample = j, (j + 1) % 2
            pupdate = p[:, jupdate::2, :]
            psample = p[:, jsample::2, :]
            qs, logf = self._stretch(pupdate, psample)
            qs = apply_bounds_batch(qs, self._period_reflect_list, self._bounds)
            qslogl, qslogp = self._evaluate(qs)
            qslogpost = self._tempered_likelihood(qslogl) + qslogp
            if hasattr(self._random, "uniform"):
                lnu = np.log(self._random.uniform(low=0.0, high=1.0, size=(self.ntemps, self.nwalkers // 2)))
            else:
                lnu = np.log(self._random.random((self.ntemps, self.nwalkers // 2)))
            with np.errstate(invalid='ignore'):
                accepts = logf + qslogpost - logpost[:, jupdate::2] > lnu
            p[:, jupdate::2, :] = np.where(accepts[:, :, None], qs, pupdate)
            logpost[:, jupdate::2] = np.where(accepts, qslogpost, logpost[:, jupdate::2])
            logl[:, jupdate::2] = np.where(accepts, qslogl, logl[:, jupdate::2])
            self.nprop[:, jupdate::2] += 1.0
            self.nprop_accepted[:, jupdate::2] += accepts
        return p, logpost, logl

    # ---- temperature swaps (ptmcmc.py:403-449) --------------------------------------------------------------------------
    def _temperature_swaps(self, betas, p, logpost, logl):
        ntemps = len(betas)
        ratios = np.zeros(ntemps - 1)
        for i in range(ntemps - 1, 0, -1):
            bi, bi1 = betas[i], betas[i - 1]
            dbeta = bi1 - bi
            iperm = self._random.permutation(self.nwalkers)
            i1perm = self._random.permutation(self.nwalkers)
            if hasattr(self._random, "uniform"):
                raccept = np.log(self._random.uniform(size=self.nwalkers))
            else:
                raccept = np.log(self._random.random(self.nwalkers))
            paccept = dbeta * (logl[i, iperm] - logl[i - 1, i1perm])
            self.nswap[i] += self.nwalkers
            self.nswap[i - 1] += self.nwalkers
            asel = (paccept > raccept)
            nacc = np.sum(asel)
            self.nswap_accepted[i] += nacc
            self.nswap_accepted[i - 1] += nacc
            ratios[i - 1] = nacc / self.nwalkers
            ptemp = np.copy(p[i, iperm[asel], :])
            logltemp = np.copy(logl[i, iperm[asel]])
            logprtemp = np.copy(logpost[i, iperm[asel]])
            p[i, iperm[asel], :] = p[i - 1, i1perm[asel], :]
            logl[i, iperm[asel]] = logl[i - 1, i1perm[asel]]
            logpost[i, iperm[asel]] = logpost[i - 1, i1perm[asel]] - dbeta * logl[i - 1, i1perm[asel]]
            p[i - 1, i1perm[asel], :] = ptemp
            logl[i - 1, i1perm[asel]] = logltemp
            logpost[i - 1, i1perm[asel]] = logprtemp + dbeta * logltemp
        return p, ratios

    def ladder_adjustment(self, time, betas0, ratios):
        """Temperature dynamics of arXiv:1501.05823 (ptmcmc.py:451-473; defined there, never called by the reference's run
        loop): returns the CHANGE of the ladder, the caller applies it."""
        betas = betas0.copy()
        decay = self.adaptation_lag / (time + self.adaptation_lag)
        kappa = decay / self.adaptation_time
        dSs = kappa * (ratios[:-1] - ratios[1:])
        deltaTs = np.diff(1 / betas[:-1])
        deltaTs *= np.exp(dSs)
        betas[1:-1] = 1 / (np.cumsum(deltaTs) + 1 / betas[0])
        return betas - betas0

    # ---- run (ptmcmc.py:338-386, 475-509) --------------------------------------------------------------------------------
    def _expand_history(self, iterations):
        shape = (self.ntemps, self.nwalkers, iterations)
        if len(self._chain_history) == 0:
            self._chain_history = np.zeros(shape + (self.dim,))
            self._logposterior_history = np.zeros(shape)
            self._loglikelihood_history = np.zeros(shape)
        else:
            self._chain_history = np.concatenate((self._chain_history, np.zeros(shape + (self.dim,))), axis=2)
            self._logposterior_history = np.concatenate((self._logposterior_history, np.zeros(shape)), axis=2)
            self._loglikelihood_history = np.concatenate((self._loglikelihood_history, np.zeros(shape)), axis=2)

    def sample(self, iterations=1, p0=None, store=True):
        if p0 is not None:
            self._p0 = np.array(p0, dtype=float).reshape((self.ntemps, self.nwalkers, self.dim))
        p = self._p0
        if p is None:
            raise ValueError("no initial positions: pass p0 [ntemps, nwalkers, dim]")
        if np.any(np.isinf(p)):
            raise ValueError("At least one parameter value was infinite.")
        if np.any(np.isnan(p)):
            raise ValueError("At least one parameter value was NaN.")
        if store:
            self._expand_history(iterations)
        if self._logposterior0 is None or self._loglikelihood0 is None:
            logl, logp = self._evaluate(p)
            logpost = self._tempered_likelihood(logl) + logp
            self._loglikelihood0, self._logposterior0 = logl, logpost
        else:
            logl, logpost = self._loglikelihood0, self._logposterior0
        for _ in range(iterations):
            p, logpost, logl = self._propose(p, logpost, logl)
            p, ratios = self._temperature_swaps(self._betas, p, logpost, logl)
            if store:
                self._chain_history[:, :, self._time, :] = p
                self._logposterior_history[:, :, self._time] = logpost
                self._loglikelihood_history[:, :, self._time] = logl
                with np.errstate(invalid='ignore', divide='ignore'):
                    self._acceptance_history.append(self.nprop_accepted / self.nprop)
                    self._tswap_acceptance_history.append(self.nswap_accepted / self.nswap)
            self._time += 1
            if self.refocus_every > 0 and self._focus_fn is not None and self._time % self.refocus_every == 0:
                cold = int(np.argmax(self._betas))                 # beta = 1
                best = int(np.argmax(np.where(np.isfinite(logl[cold]), logl[cold], -np.inf)))
                self._focus_fn(np.array(p[cold, best], dtype=np.float64))
                self.n_refocus += 1
        self._p0, self._logposterior0, self._loglikelihood0 = p, logpost, logl
        return p, logpost, logl

    @property
    def betas(self):
        return self._betas

    @property
    def acceptance_fraction(self):
        with np.errstate(invalid='ignore', divide='ignore'):
            return self.nprop_accepted / self.nprop

    @property
    def tswap_acceptance_fraction(self):
        with np.errstate(invalid='ignore', divide='ignore'):
            return self.nswap_accepted / self.nswap
