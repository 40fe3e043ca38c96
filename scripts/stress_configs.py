#!/usr/bin/env python
"""GPU stress run over data configurations (segment length, band, detector subsets): 20 011 walkers per case, production
kernel against the all-FP64 validation kernel."""
import sys

import numpy as np

sys.path.insert(0, ".")
from bajes_b200 import GWLikelihood, synthetic as syn   # noqa: E402


def tol(ref):
    return 1e-6 * np.abs(ref) + 1e-4


W = 20011
base = {"bns": syn.CONFIGS["C2_small"], "bbh": syn.CONFIGS["C1"]}
cases = [("bns", dict(seglen=4.0)), ("bns", dict(seglen=64.0)), ("bns", dict(seglen=256.0, f_max=1024.0, srate=2048.0)),
         ("bns", dict(f_min=10.0)), ("bns", dict(f_min=30.0, f_max=511.0)), ("bns", dict(ifos=["H1"])),
         ("bns", dict(ifos=["L1", "V1"], seglen=32.0)), ("bns", dict(f_min=23.3, f_max=1777.7)),
         ("bbh", dict(seglen=8.0, f_min=10.0)), ("bbh", dict(seglen=2.0, f_max=512.0, srate=1024.0)),
         ("bbh", dict(ifos=["H1", "L1", "V1"], seglen=16.0))]
for kind, upd in cases:
    cfg = dict(base[kind])
    cfg.update(upd)
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    Xa, names = syn.draw_walkers(cfg, W // 2, 51, "narrow")
    Xb, _ = syn.draw_walkers(cfg, W - W // 2, 52, "wide")
    X = np.concatenate([Xa, Xb])
    const = syn.constants(cfg)
    b = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"], sincos="fp64").log_like_batch(X, names, const)
    fast = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"])
    a = fast.log_like_batch(X, names, const)
    fin = np.isfinite(b)
    r = np.abs(a[fin] - b[fin]) / tol(b[fin])
    print("%-4s %-58s bins=%7d finite=%5d pattern_ok=%s max|d|/tol=%.3f 5th=%.3f p99.9=%.3f"
          % (kind, upd, fast.engine.n_bins, fin.sum(), np.array_equal(np.isfinite(a), fin), r.max(), np.sort(r)[-5],
             np.quantile(r, 0.999)), flush=True)
