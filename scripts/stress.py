#!/usr/bin/env python
"""Ad-hoc GPU stress run: large and odd batch sizes, production kernel against the FP64 validation kernel."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bajes_b200 import GWLikelihood, synthetic as syn   # noqa: E402


def tol(ref):
    return 1e-6 * np.abs(ref) + 1e-4


for name, W in (("C2_small", 100003), ("C2", 20001), ("C1", 262145)):
    cfg = syn.CONFIGS[name]
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    fast = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"])
    slow = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"], sincos="fp64")
    Xa, names = syn.draw_walkers(cfg, W // 2, 11, "narrow")
    Xb, _ = syn.draw_walkers(cfg, W - W // 2, 12, "wide")
    X = np.concatenate([Xa, Xb])
    const = syn.constants(cfg)
    t0 = time.time()
    a = fast.log_like_batch(X, names, const)
    t1 = time.time()
    b = slow.log_like_batch(X, names, const)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    r = np.abs(a[fin] - b[fin]) / tol(b[fin])
    print("%-9s W=%6d finite=%6d  max|d|/tol=%.3f  p99=%.3f  fast call %.1f ms  kernel %.2f ms" %
          (name, W, fin.sum(), r.max(), np.quantile(r, 0.99), (t1 - t0) * 1e3, fast.engine.last_kernel_ms()[1]))
    assert r.max() <= 1.0
