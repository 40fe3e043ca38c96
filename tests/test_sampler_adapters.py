"""
N1: the pool-shaped adapter under an UNMODIFIED bajes sampler.  bajes' parallel-tempered
ensemble sampler (bajes/inf/sampler/ptmcmc.py:182-470, the only sampler of the reference that runs offline) evaluates
every half-ensemble with ``pool.map(posterior.log_likeprior, points)`` (:272-281); with ``BatchedPool`` and the batched
``Posterior`` that becomes one batched likelihood call per half-ensemble, and the chain must not change by a bit.
Runs only where the reference tree exists (this build container); the likelihood is a CPU stand-in with the product's
batched interface, so no GPU is needed: what is under test is the adapter contract, not the kernel.
"""
import numpy as np
import pytest

from bajes_b200 import Posterior
from bajes_b200.pool import BatchedPool
from oracle import shims

pytestmark = pytest.mark.needs_reference

MU = np.array([0.3, -0.2, 1.0])
SIG = np.array([0.5, 1.0, 0.2])


class _Like(object):
    """log_like(dict) as bajes' Likelihood, log_like_batch(X, names, const) as bajes_b200.GWLikelihood."""
    names = ("a", "b", "c")

    def log_like(self, p):
        x = np.array([p[k] for k in self.names])
        return float(self.log_like_batch(x[None], list(self.names), {})[0])

    def log_like_batch(self, X, names, const):
        idx = [names.index(k) for k in self.names]
        t = ((np.asarray(X)[:, idx] - MU) / SIG) ** 2
        return -0.5 * (t[:, 0] + t[:, 1] + t[:, 2])


def _run(batched, seed=7, iters=15):
    from bajes.inf.likelihood import Posterior as ReferencePosterior
    from bajes.inf.prior import Constant, Parameter, Prior
    from bajes.inf.sampler.ptmcmc import _PTMCMC, BajesPTMCMCProposal
    prior = Prior([Parameter(name="a", min=-3, max=3), Parameter(name="b", min=-3, max=3, periodic=1),
                   Parameter(name="c", min=-3, max=3)], constants=[Constant("off", 1.5)])
    like = _Like()
    rng = np.random.RandomState(seed)
    np.random.seed(seed)
    pool = None
    if batched:
        post = Posterior(like, prior)
        pool = BatchedPool(post, processes=4)
    else:
        post = ReferencePosterior(like, prior)
    sampler = _PTMCMC(32, post, BajesPTMCMCProposal(prior, props={"str": 1.0}), ntemps=3, random=rng, pool=pool)
    sampler._p0 = rng.uniform(-1, 1, size=(3, 32, 3))
    sampler._expand_history(iterations=iters)
    sampler.sample(iterations=iters)
    return sampler, pool


def test_reference_ptmcmc_through_the_batched_pool_is_bit_identical():
    if not shims.reference_available():
        pytest.skip("reference tree not present")
    shims.import_reference()
    ref, _ = _run(False)
    got, pool = _run(True)
    assert np.array_equal(ref._chain_history, got._chain_history)
    assert np.array_equal(ref._loglikelihood_history, got._loglikelihood_history)
    assert np.array_equal(ref._logposterior_history, got._logposterior_history)
    # 1 initial evaluation + 2 half-ensembles x 15 iterations, each ONE batched call over 3 temperatures x 16 walkers
    assert pool.n_batched_calls == 31 and pool.n_points == 3 * 32 + 15 * 2 * 3 * 16
    assert np.isfinite(got._loglikelihood_history).all() and got._time == 15
