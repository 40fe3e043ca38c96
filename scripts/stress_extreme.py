#!/usr/bin/env python
"""GPU stress run over an EXTREME parameter box (sub-solar to heavy masses, mass ratios to 8, spins to 0.9, tides to 5000,
1 Mpc to 1 Gpc, +-0.1 s time shifts): production kernels against the all-FP64 validation kernel."""
import sys

import numpy as np

sys.path.insert(0, ".")
from bajes_b200 import GWLikelihood, synthetic as syn   # noqa: E402


def tol(ref):
    return 1e-6 * np.abs(ref) + 1e-4


W = 100003
rng = np.random.default_rng(5)
for name, mc_lo, mc_hi in (("C2_small", 0.25, 3.0), ("C1", 3.0, 120.0)):
    cfg = syn.CONFIGS[name]
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    _, names = syn.draw_walkers(cfg, 2, 1, "wide")
    cols = dict(mchirp=np.exp(rng.uniform(np.log(mc_lo), np.log(mc_hi), W)), q=rng.uniform(1, 8, W),
                s1z=rng.uniform(-0.9, 0.9, W), s2z=rng.uniform(-0.9, 0.9, W), lambda1=rng.uniform(0, 5000, W),
                lambda2=rng.uniform(0, 5000, W), distance=np.exp(rng.uniform(0, np.log(1000), W)),
                cos_iota=rng.uniform(-1, 1, W), ra=rng.uniform(0, 2 * np.pi, W), dec=np.arcsin(rng.uniform(-1, 1, W)),
                psi=rng.uniform(0, np.pi, W), time_shift=rng.uniform(-0.1, 0.1, W), phi_ref=rng.uniform(0, 2 * np.pi, W))
    if name == "C1":
        cols["lambda1"][:] = 0.0
        cols["lambda2"][:] = 0.0
    X = np.stack([cols[k] for k in names], axis=1)
    const = syn.constants(cfg)
    for approx in (cfg["approx"], "TaylorF2_5.5PN_7.5PNTides" if name != "C1" else "TaylorF2_5.5PN"):
        b = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos="fp64").log_like_batch(X, names, const)
        for sc in ("mufu", "poly32"):
            a = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos=sc).log_like_batch(X, names, const)
            fin = np.isfinite(b)
            ok = np.array_equal(np.isfinite(a), fin)
            r = np.abs(a[fin] - b[fin]) / tol(b[fin])
            i = np.flatnonzero(fin)[np.argmax(r)]
            print("%-9s %-28s %-6s finite=%6d pattern_ok=%s max|d|/tol=%.3f 5th=%.3f p99.9=%.3f  worst: logL=%.4g %s"
                  % (name, approx, sc, fin.sum(), ok, r.max(), np.sort(r)[-5], np.quantile(r, 0.999), b[i],
                     {k: float("%.4g" % v) for k, v in zip(names, X[i])}), flush=True)
