#!/usr/bin/env python
"""
BUILD CONTAINER ONLY (needs the read-only reference tree).  Runs the UNMODIFIED parallel-tempered ensemble sampler of the
reference (bajes/inf/sampler/ptmcmc.py:182-470 -- the only bajes sampler that runs offline) on a BNS injection, with a
bajes ``Prior`` and the batched ``Posterior`` / ``BatchedPool`` adapters of this repo, and records EVERY batch the sampler
sent through ``pool.map(posterior.log_likeprior, points)`` (:272-281) together with the values it got back.

The likelihood behind the adapter is the oracle port (CPU: no GPU exists here); the GPU box has no reference tree, so the
two halves meet in ``tests/golden/ptmcmc_trace.npz``: ``tests/test_gpu_parity.py::test_reference_ptmcmc_call_trace_on_the_gpu``
replays the recorded call sequence through the same adapters on the GPU likelihood and checks every value the
sampler saw.  Config: the reduced C2 (16 s, H1/L1/V1, 32 449 bins), 2 temperatures x 16 walkers x 12 iterations.

    python scripts/make_ptmcmc_trace.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bajes_b200 import Posterior, synthetic as syn      # noqa: E402
from bajes_b200.pool import BatchedPool                  # noqa: E402
from oracle import harness, shims                        # noqa: E402

NTEMPS, NWALKERS, ITERS, SEED = 2, 16, 12, 11


class RecordingLikelihood(object):
    """log_like_batch(X, names, const) of the product interface on top of the oracle port; records every call"""
    logZ_noise = 0.0

    def __init__(self, port_like):
        self.port = port_like
        self.calls = []

    def log_like(self, p):
        return self.port.log_like(p)

    def log_like_batch(self, X, names, const=None):
        out = harness.eval_batch(self.port, X, names, const or {})
        self.calls.append((np.array(X), list(names), out.copy()))
        return out


def main():
    shims.import_reference()
    from bajes.inf.prior import Constant, Parameter, Prior
    from bajes.inf.sampler.ptmcmc import _PTMCMC, BajesPTMCMCProposal
    cfg = syn.CONFIGS["C2_small"]
    cl = harness.port_classes()
    net = syn.build_network(cfg, cl)
    port_like = harness.port_likelihood(cfg, net=net)
    names = syn.WALKER_NAMES
    p0, _ = syn.draw_walkers(cfg, 4096, 123, "narrow")
    lo, hi = p0.min(axis=0), p0.max(axis=0)
    pad = 0.5 * (hi - lo)
    bounds = np.stack([lo - pad, hi + pad], axis=1)
    bounds[names.index("q"), 0] = max(1.0, bounds[names.index("q"), 0])
    bounds[names.index("cos_iota")] = np.clip(bounds[names.index("cos_iota")], -1, 1)
    for k in ("lambda1", "lambda2"):
        bounds[names.index(k), 0] = max(0.0, bounds[names.index(k), 0])
    periodic = {"psi": 1, "phi_ref": 1}
    const = syn.constants(cfg)
    prior = Prior([Parameter(name=n, min=float(b[0]), max=float(b[1]), periodic=periodic.get(n, 0))
                   for n, b in zip(names, bounds)],
                  constants=[Constant(k, v) for k, v in const.items()])
    like = RecordingLikelihood(port_like)
    post = Posterior(like, prior)
    pool = BatchedPool(post, processes=1)
    rng = np.random.RandomState(SEED)
    np.random.seed(SEED)
    sampler = _PTMCMC(NWALKERS, post, BajesPTMCMCProposal(prior, props={"str": 1.0}), ntemps=NTEMPS, random=rng, pool=pool)
    start, _ = syn.draw_walkers(cfg, NTEMPS * NWALKERS, 321, "narrow")
    sampler._p0 = start.reshape(NTEMPS, NWALKERS, len(names))
    sampler._expand_history(iterations=ITERS)
    sampler.sample(iterations=ITERS)
    X = np.concatenate([c[0] for c in like.calls])
    sizes = np.array([len(c[0]) for c in like.calls])
    logl = np.concatenate([c[2] for c in like.calls])
    assert all(c[1] == like.calls[0][1] for c in like.calls)
    dets = net[2]
    out = os.path.join(ROOT, "tests", "golden", "ptmcmc_trace.npz")
    np.savez_compressed(out, X=X, sizes=sizes, logl=logl, names=np.array(like.calls[0][1]),
                        const_keys=np.array(sorted(const)), const_vals=np.array([const[k] for k in sorted(const)]),
                        chain=np.array(sampler._chain_history), loglike_hist=np.array(sampler._loglikelihood_history),
                        gmst_reference=np.array([dets[i].gmst_reference for i in cfg["ifos"]]),
                        sday=np.array([dets[i].sday for i in cfg["ifos"]]),
                        shape=np.array([NTEMPS, NWALKERS, ITERS, SEED]))
    print("ptmcmc_trace.npz: %d batched calls, %d points, batch sizes %s..., logL in [%.1f, %.1f], acceptance %.2f"
          % (len(sizes), len(X), sizes[:4].tolist(), np.nanmin(logl[np.isfinite(logl)]), np.nanmax(logl[np.isfinite(logl)]),
             float(np.mean(sampler.nprop_accepted / np.maximum(sampler.nprop, 1)))))


if __name__ == "__main__":
    main()
