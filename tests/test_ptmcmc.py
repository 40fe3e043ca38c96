"""
``bajes_b200.ptmcmc.PTEnsembleSampler`` -- the batch-native restatement of bajes' parallel-tempered ensemble kernel
(bajes/inf/sampler/ptmcmc.py:184-470 with the stretch move of proposal.py:92-117) -- pinned to the reference:

  * against the UNMODIFIED ``_PTMCMC`` where the reference tree exists (same prior object, same seeds): chains, lnL and
    ln-posterior histories bit for bit;
  * everywhere, against ``tests/golden/ptmcmc_trace.npz`` -- every batch the unmodified sampler sent to the likelihood on the
    reduced C2 injection, the values it got back and the chain it produced (scripts/make_ptmcmc_trace.py): the restatement
    must ask for exactly the same points in the same order and end with the same chain;
  * on the GPU (``-m gpu``): the same run with the CUDA likelihood behind it.
"""
import os

import numpy as np
import pytest

from bajes_b200 import Posterior, synthetic as syn
from bajes_b200.ensemble import BoxPrior
from bajes_b200.ptmcmc import PTEnsembleSampler, apply_bounds_batch, default_beta_ladder
from oracle import shims

from conftest import GOLDEN, load_golden

MU = np.array([0.3, -0.2, 1.0])
SIG = np.array([0.5, 1.0, 0.2])


class _Gauss(object):
    names = ("a", "b", "c")

    def log_like(self, p):
        x = np.array([p[k] for k in self.names])
        return float(self.log_like_batch(x[None], list(self.names), {})[0])

    def log_like_batch(self, X, names, const):
        idx = [names.index(k) for k in self.names]
        t = ((np.asarray(X)[:, idx] - MU) / SIG) ** 2
        return -0.5 * (t[:, 0] + t[:, 1] + t[:, 2])


@pytest.mark.needs_reference
def test_bit_identical_to_the_unmodified_reference_sampler():
    if not shims.reference_available():
        pytest.skip("reference tree not present")
    shims.import_reference()
    from bajes.inf.likelihood import Posterior as ReferencePosterior
    from bajes.inf.prior import Constant, Parameter, Prior
    from bajes.inf.sampler.ptmcmc import _PTMCMC, BajesPTMCMCProposal, default_beta_ladder as ref_ladder
    prior = Prior([Parameter(name="a", min=-3, max=3), Parameter(name="b", min=-1.2, max=1.2, periodic=1),
                   Parameter(name="c", min=0.4, max=1.6)], constants=[Constant("off", 1.5)])
    like = _Gauss()
    seed, iters, ntemps, nw = 5, 25, 4, 24
    p0 = np.random.RandomState(99).uniform(0.5, 1.1, size=(ntemps, nw, 3))

    rng = np.random.RandomState(seed)
    np.random.seed(seed)
    ref = _PTMCMC(nw, ReferencePosterior(like, prior), BajesPTMCMCProposal(prior, props={"str": 1.0}), ntemps=ntemps, random=rng)
    ref._p0 = p0.copy()
    ref._expand_history(iterations=iters)
    ref.sample(iterations=iters)

    rng = np.random.RandomState(seed)
    np.random.seed(seed)
    got = PTEnsembleSampler(nw, Posterior(like, prior), ntemps=ntemps, random=rng, rng="reference")
    got.sample(iterations=iters, p0=p0.copy())

    assert np.array_equal(ref._betas, got.betas)
    assert np.array_equal(ref._chain_history, got._chain_history)
    assert np.array_equal(ref._loglikelihood_history, got._loglikelihood_history)
    assert np.array_equal(ref._logposterior_history, got._logposterior_history)
    assert np.array_equal(ref.nprop_accepted, got.nprop_accepted) and np.array_equal(ref.nswap_accepted, got.nswap_accepted)
    # the walls were hit (reflection and periodic wrap exercised) and swaps happened
    assert got.nswap_accepted.sum() > 0 and 0.05 < got.acceptance_fraction.mean() < 0.95
    assert got.n_batched_calls == 1 + 2 * iters and got.n_points == ntemps * nw * (1 + iters)
    for nd, nt, tm in ((3, 4, None), (13, None, 50.0), (13, 5, np.inf), (120, 3, None)):
        assert np.array_equal(ref_ladder(nd, ntemps=nt, Tmax=tm), default_beta_ladder(nd, ntemps=nt, Tmax=tm))


def test_boundary_moves_match_the_reference_expressions():
    bounds = [[-1.0, 2.0], [0.0, np.pi], [5.0, 6.0]]
    per = [0, 1, 0]
    Q = np.array([[-1.5, -0.4, 5.5], [2.75, 3.5, 6.25], [0.0, 1.0, 5.0], [-7.0, 11.0, 4.0]])
    out = apply_bounds_batch(Q, per, bounds)
    # utils.py:178-198 written out
    want = Q.copy()
    for r in range(len(Q)):
        for i, (lo, up) in enumerate(bounds):
            x = Q[r, i]
            if per[i]:
                want[r, i] = up - np.abs(x - lo) % (up - lo) if x < lo else ((x - lo) % (up - lo) + lo if x > up else x)
            else:
                want[r, i] = 2. * lo - x if x < lo else (2. * up - x if x > up else x)
    assert np.array_equal(out, want)
    assert np.array_equal(out[2], Q[2])           # in bounds: untouched


def _trace_setup():
    """prior, start and seeds of scripts/make_ptmcmc_trace.py, rebuilt without the reference"""
    gold = load_golden("ptmcmc_trace.npz")
    cfg = syn.CONFIGS["C2_small"]
    names = [str(n) for n in gold["names"]]
    assert names == syn.WALKER_NAMES
    p0, _ = syn.draw_walkers(cfg, 4096, 123, "narrow")
    lo, hi = p0.min(axis=0), p0.max(axis=0)
    pad = 0.5 * (hi - lo)
    bounds = np.stack([lo - pad, hi + pad], axis=1)
    bounds[names.index("q"), 0] = max(1.0, bounds[names.index("q"), 0])
    bounds[names.index("cos_iota")] = np.clip(bounds[names.index("cos_iota")], -1, 1)
    for k in ("lambda1", "lambda2"):
        bounds[names.index(k), 0] = max(0.0, bounds[names.index(k), 0])
    const = {str(k): float(v) for k, v in zip(gold["const_keys"], gold["const_vals"])}
    prior = BoxPrior(names, bounds.tolist(), const, periodic=[1 if n in ("psi", "phi_ref") else 0 for n in names])
    ntemps, nwalkers, iters, seed = (int(v) for v in gold["shape"])
    start, _ = syn.draw_walkers(cfg, ntemps * nwalkers, 321, "narrow")
    return gold, cfg, names, const, prior, ntemps, nwalkers, iters, seed, start.reshape(ntemps, nwalkers, len(names))


class _Replay(object):
    """likelihood that answers from the recorded trace and insists on being asked exactly what the reference asked"""
    logZ_noise = 0.0

    def __init__(self, gold):
        self.X, self.logl, self.sizes = gold["X"], gold["logl"], gold["sizes"]
        self.call, self.row = 0, 0

    def log_like_batch(self, X, names, const=None):
        n = len(X)
        assert self.call < len(self.sizes) and n == self.sizes[self.call], "batch %d: %d points, the reference sent %d" % (
            self.call, n, self.sizes[self.call] if self.call < len(self.sizes) else -1)
        assert np.array_equal(np.asarray(X), self.X[self.row:self.row + n]), "batch %d differs from the reference's" % self.call
        out = self.logl[self.row:self.row + n].copy()
        self.call += 1
        self.row += n
        return out


def test_asks_for_the_same_points_as_the_reference_and_ends_with_its_chain():
    gold, cfg, names, const, prior, ntemps, nwalkers, iters, seed, start = _trace_setup()
    like = _Replay(gold)
    rng = np.random.RandomState(seed)
    np.random.seed(seed)
    s = PTEnsembleSampler(nwalkers, Posterior(like, prior), ntemps=ntemps, random=rng, rng="reference")
    s.sample(iterations=iters, p0=start.copy())
    assert like.call == len(gold["sizes"]) and like.row == len(gold["X"])
    assert np.array_equal(s._chain_history, gold["chain"])
    assert np.array_equal(s._loglikelihood_history, gold["loglike_hist"])


def test_vector_mode_samples_a_gaussian():
    """the array-drawn mode is a different stream, not a different sampler: moments of a tempered Gaussian target"""
    prior = BoxPrior(["a", "b", "c"], [[-6, 6], [-6, 6], [-6, 6]], {})
    s = PTEnsembleSampler(200, Posterior(_Gauss(), prior), ntemps=3, random=np.random.default_rng(3), rng="vector")
    p0 = np.random.default_rng(4).uniform(-1, 1, size=(3, 200, 3))
    s.sample(iterations=300, p0=p0)
    cold = s._chain_history[0, :, 100:, :].reshape(-1, 3)
    assert np.all(np.abs(cold.mean(axis=0) - MU) < 0.05)
    assert np.all(np.abs(cold.std(axis=0) / SIG - 1.0) < 0.08)
    hot = s._chain_history[-1, :, 100:, :].reshape(-1, 3)
    assert np.all(hot.std(axis=0) > 1.3 * cold.std(axis=0))            # beta < 1 broadens the target
    assert s.n_batched_calls == 1 + 2 * 300


@pytest.mark.gpu
def test_reference_chain_on_the_gpu_likelihood():
    """the recorded reference run, re-run with the CUDA likelihood behind the sampler: same chain"""
    from bajes_b200 import GWLikelihood
    from conftest import port_network, product_network_from_port
    gold, cfg, names, const, prior, ntemps, nwalkers, iters, seed, start = _trace_setup()
    ifos, datas, dets, noises, f = product_network_from_port(cfg, port_network(cfg, gold))
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"])
    rng = np.random.RandomState(seed)
    np.random.seed(seed)
    s = PTEnsembleSampler(nwalkers, Posterior(like, prior), ntemps=ntemps, random=rng, rng="reference")
    s.sample(iterations=iters, p0=start.copy())
    same = np.all(s._chain_history == gold["chain"], axis=-1)
    assert same.mean() == 1.0, "chain departs from the reference's in %d of %d states" % ((~same).sum(), same.size)
    ref = gold["loglike_hist"]
    assert np.all(np.abs(s._loglikelihood_history - ref) <= 1e-6 * np.abs(ref) + 1e-4)


def test_refocus_hook_hands_the_best_cold_point_to_the_likelihood_and_moves_no_random_number():
    class _G(_Gauss):
        def __init__(self):
            self.focused = []

        def focus(self, row, names=None, const=None):
            assert list(names) == ["a", "b", "c"] and const == {}
            self.focused.append(np.array(row))

    prior = BoxPrior(["a", "b", "c"], [[-6, 6], [-6, 6], [-6, 6]], {})
    p0 = np.random.default_rng(4).uniform(-1, 1, size=(3, 40, 3))
    plain = PTEnsembleSampler(40, Posterior(_Gauss(), prior), ntemps=3, random=np.random.default_rng(3), rng="vector")
    plain.sample(iterations=30, p0=p0.copy())
    like = _G()
    s = PTEnsembleSampler(40, Posterior(like, prior), ntemps=3, random=np.random.default_rng(3), rng="vector", refocus_every=10)
    s.sample(iterations=30, p0=p0.copy())
    assert s.n_refocus == 3 and len(like.focused) == 3
    assert np.array_equal(s._chain_history, plain._chain_history)          # the hook draws nothing
    # what it handed over is the best point of the cold chain at that iteration
    for k, it in enumerate((9, 19, 29)):
        j = int(np.argmax(s._loglikelihood_history[0, :, it]))
        assert np.array_equal(like.focused[k], s._chain_history[0, j, it])
