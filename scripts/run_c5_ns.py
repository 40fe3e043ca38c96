#!/usr/bin/env python
"""
BASELINE config 5, nested-sampling flavour: a static nested sampler on the 128 s BNS injection whose replacement points
come from the lock-step random walk (``bajes_b200.rwalk.LockstepRandomWalk``, the batched stand-in for bajes'
``BajesDynestyProposal.sample_rwalk``, dynesty.py:503-614) through ``BatchedPool.map(walk.propose, args)`` -- the call
dynesty makes with its queue of ``evolve_point`` arguments (dynesty itself is not installable offline).

    python scripts/run_c5_ns.py --nlive 1024 --queue 256 --rounds 12

Every proposal round of the walk is ONE batched GPU likelihood call over all still-active chains.  Prints one JSON
line: likelihood evaluations, batched calls, mean batch, wall time, evaluations per second end to end (host -> device
-> host every call), the running evidence estimate.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--nlive", type=int, default=1024)
    ap.add_argument("--queue", type=int, default=256)
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--walks", type=int, default=20)
    ap.add_argument("--maxmcmc", type=int, default=200)
    args = ap.parse_args()

    from bajes_b200 import GWLikelihood, Posterior, synthetic as syn
    from bajes_b200.ensemble import BoxPrior
    from bajes_b200.pool import BatchedPool
    from bajes_b200.rwalk import LockstepRandomWalk

    cfg = syn.CONFIGS[args.config]
    cl = syn.product_classes()
    like = GWLikelihood(*syn.build_network(cfg, cl), cfg["srate"], cfg["seglen"], cfg["approx"])
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    p0, _ = syn.draw_walkers(cfg, 4096, 123, "narrow")
    lo, hi = p0.min(axis=0), p0.max(axis=0)
    periodic = [1 if n in ("ra", "psi", "phi_ref") else 0 for n in names]
    prior = BoxPrior(names, np.stack([lo, hi], axis=1), const, periodic=periodic)
    post = Posterior(like, prior)
    pool = BatchedPool(post, processes=args.queue + 1)
    walk = LockstepRandomWalk(args.maxmcmc, walks=args.walks, nact=2.0, posterior=post, seed=3)
    n = len(names)
    kwargs = dict(periodic=[i for i, p in enumerate(periodic) if p] or None,
                  reflective=[i for i, p in enumerate(periodic) if not p] or None, nonbounded=np.zeros(n, dtype=bool))

    rng = np.random.default_rng(11)
    t0 = time.perf_counter()
    U = rng.uniform(size=(args.nlive, n))
    V = post.prior_transform_batch(U)
    L = np.asarray(pool.map(post.log_like, V))                       # ONE batched call for the initial live points
    n_eval0 = len(U)
    logz, logx, it = -np.inf, 0.0, 0
    replaced = 0
    for _ in range(args.rounds):
        order = np.argsort(L)
        loglstar = float(L[order[0]])
        # proposal scale from the live points, start points drawn among those above the current bound
        axes = np.linalg.cholesky(np.cov(U.T) + 1e-12 * np.eye(n))
        starts = rng.choice(order[1:], size=args.queue)
        qargs = [(U[s].copy(), loglstar, axes, 1.0, post.prior_transform, post.log_like, None, dict(kwargs))
                 for s in starts]
        new = pool.map(walk.propose, qargs)                          # dynesty: self.M(evolve_point, args)
        for u_new, v_new, l_new, ncall, blob in new:
            worst = int(np.argmin(L))
            if l_new <= L[worst]:
                continue                                             # (dynesty re-queues such a point; skipped here)
            # standard static-NS bookkeeping: the prior volume shrinks by nlive / (nlive + 1) per replaced point
            logw = logx + np.log(1.0 - np.exp(-1.0 / args.nlive))
            logz = np.logaddexp(logz, L[worst] + logw)
            logx -= 1.0 / args.nlive
            U[worst], V[worst], L[worst] = u_new, v_new, l_new
            replaced += 1
            it += 1
    wall = time.perf_counter() - t0
    n_eval = n_eval0 + walk.n_evaluations
    print(json.dumps({
        "config": args.config, "n_bins": like.engine.n_bins, "n_ifo": len(like.ifos), "nlive": args.nlive,
        "queue": args.queue, "rounds": args.rounds, "walks": args.walks, "maxmcmc": args.maxmcmc,
        "replaced_points": replaced, "likelihood_evaluations": n_eval, "batched_calls": 1 + walk.n_batched_calls,
        "mean_batch": walk.n_evaluations / max(1, walk.n_batched_calls), "wall_s": wall,
        "evals_per_s_end_to_end": n_eval / wall, "logL_bound": float(np.min(L)), "logL_max": float(np.max(L)),
        "logZ_so_far": float(logz), "log_prior_volume": float(logx)}))


if __name__ == "__main__":
    main()
