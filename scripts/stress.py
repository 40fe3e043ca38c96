#!/usr/bin/env python
"""GPU stress run: 100 003 walkers per case, production kernels against the all-FP64 validation kernel."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bajes_b200 import GWLikelihood, synthetic as syn   # noqa: E402


def tol(ref):
    return 1e-6 * np.abs(ref) + 1e-4


W = int(sys.argv[1]) if len(sys.argv) > 1 else 100003
cases = []
for approx in ("TaylorF2_3.5PN", "TaylorF2_5.5PN", "TaylorF2_5.5PN_7.5PNTides", "TaylorF2_5.5PN_3.5PNQM_7.5PNTides",
               "TaylorF2_5.5PN_7.5PNTides2020"):
    cases.append(("C2_small", approx, {}, 0, 0))
cases.append(("C2_small", "TaylorF2_3.5PN", dict(marg_phi_ref=True), 0, 0))
cases.append(("C2_small", "TaylorF2_3.5PN", {}, 6, 0))
cases.append(("C2_small", "TaylorF2_3.5PN", {}, 6, 4))
cases.append(("C1", "TaylorF2_3.5PN", dict(marg_phi_ref=True), 0, 0))
cases.append(("C1", "TaylorF2_5.5PN", {}, 0, 0))
cases.append(("C1", "TaylorF2_3.5PN", {}, 5, 3))
worst = 0.0
for name, approx, kw, nsp, nw in cases:
    cfg = syn.CONFIGS[name]
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    ex = syn.extras_setup(cfg, nsp, nw) if (nsp or nw) else {}
    Xa, names = syn.draw_walkers(cfg, W // 2, 11, "narrow")
    Xb, _ = syn.draw_walkers(cfg, W - W // 2, 12, "wide")
    X = np.concatenate([Xa, Xb])
    if nsp or nw:
        Xe, ne = syn.extras_columns(cfg, W, 13, nsp, nw)
        X, names = np.ascontiguousarray(np.concatenate([X, Xe], axis=1)), names + ne
    const = syn.constants(cfg)
    b = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos="fp64", **kw, **ex).log_like_batch(X, names, const)
    for sc in ("mufu", "poly32"):
        t0 = time.time()
        a = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos=sc, **kw, **ex).log_like_batch(X, names, const)
        fin = np.isfinite(b)
        same = np.array_equal(np.isfinite(a), fin)
        r = np.abs(a[fin] - b[fin]) / tol(b[fin])
        worst = max(worst, float(r.max()))
        print("%-9s %-34s %-22s spcal=%d weights=%d %-6s finite=%6d pattern_ok=%s max|d|/tol=%.3f 5th=%.3f p99.9=%.3f"
              % (name, approx, kw, nsp, nw, sc, fin.sum(), same, r.max(), np.sort(r)[-5], np.quantile(r, 0.999)), flush=True)
print("worst", worst)
