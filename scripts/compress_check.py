#!/usr/bin/env python
"""
Compressed frequency axis against the full grid, on the GPU:

  * node counts of the context (bj_compress_info),
  * logL of W walkers (narrow + wide boxes) on the compressed axis, on the full grid (compress=False) and with the all-FP64
    validation kernel: worst |dlogL| / tol of each against FP64,
  * walkers outside the design (time shifts beyond tau_max, sub-design chirp masses): flagged, evaluated on the full grid,
  * kernel times (CUDA events inside the library) for W walkers and for 64.

  python scripts/compress_check.py [--config C2] [--walkers 4096] [--approx TaylorF2_3.5PN]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bajes_b200 import GWLikelihood, synthetic as syn  # noqa: E402


def tol(ref):
    return 1e-6 * np.abs(np.where(np.isfinite(ref), ref, 0.0)) + 1e-4


def ratio(got, ref):
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), "finite / -inf pattern differs"
    return float(np.max(np.abs(got[fin] - ref[fin]) / tol(ref[fin]))) if fin.any() else 0.0


def timed(like, X, names, const, reps=5):
    like.log_like_batch(X, names, const)
    ms = []
    for _ in range(reps):
        like.log_like_batch(X, names, const)
        ms.append(like.engine.last_kernel_ms())
    return np.median(np.asarray(ms), axis=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--approx", default="")
    a = ap.parse_args()
    cfg = dict(syn.CONFIGS[a.config])
    approx = a.approx or cfg["approx"]
    cl = syn.product_classes()
    t0 = time.time()
    net = syn.build_network(cfg, cl)
    const = syn.constants(cfg)
    print("# %s %s: network built in %.1f s" % (a.config, approx, time.time() - t0), flush=True)
    t0 = time.time()
    comp = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx)
    t1 = time.time()
    full = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, compress=False)
    t2 = time.time()
    ref = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos="fp64")
    info = comp.engine.compress_info()
    info0_nodes = info["nodes"]
    print("context: compressed %.2f s, full grid %.2f s" % (t1 - t0, t2 - t1), flush=True)
    for lv in range(2):
        info = comp.engine.compress_info(lv)
        if info["nodes"]:
            print("level %d (mchirp >= %.2f, |tau| <= %.2f s): %d bins -> %d nodes (x%.1f), %d of them plain bins"
                  % (lv, info["mchirp_min"], info["tau_max"], info["n_bins"], info["nodes"], info["n_bins"] / info["nodes"],
                     info["kept_bins"]))
    W = a.walkers
    Xn, names = syn.draw_walkers(cfg, W // 2, 31, "narrow")
    Xw, _ = syn.draw_walkers(cfg, W - W // 2, 32, "wide")
    X = np.concatenate([Xn, Xw])
    r = ref.log_like_batch(X, names, const)
    g = comp.log_like_batch(X, names, const)
    steep = comp.engine.last_level_counts()
    f = full.log_like_batch(X, names, const)
    print("in-design batch (%d walkers): compressed vs fp64 %.3f tol, full grid vs fp64 %.3f tol, compressed vs full grid "
          "%.3f tol; walkers per table (level 0, level 1, full grid) %s; finite %d" % (W, ratio(g, r), ratio(f, r), ratio(g, f), steep,
                                                              int(np.isfinite(r).sum())), flush=True)
    # out-of-design walkers: large time shifts, light systems
    Xs = X[: min(W, 512)].copy()
    it, im = names.index("time_shift"), names.index("mchirp")
    Xs[0::4, it] = 0.3
    Xs[1::4, it] = -0.18
    Xs[2::4, im] = 0.3
    r2 = ref.log_like_batch(Xs, names, const)
    g2 = comp.log_like_batch(Xs, names, const)
    print("out-of-design batch (%d walkers, 3/4 of them outside level 0): compressed-context vs fp64 %.3f tol; walkers per "
          "table %s" % (len(Xs), ratio(g2, r2), comp.engine.last_level_counts()), flush=True)
    mo = timed(comp, Xs, names, const)
    print("         kernels %.4f ms for that batch" % mo[1])
    if info0_nodes:
        foc = comp.focus(syn.injection_params(cfg))
        gf = comp.log_like_batch(X, names, const)
        cnt = comp.engine.last_level_counts()
        lv = comp.engine.last_levels(len(X))
        print("focus table around the injection: %d nodes in %d runs; walkers per table (level 0, level 1, full grid, focus) %s; "
              "vs fp64: focused walkers %.4f tol, all %.3f tol" % (foc["nodes"], foc["blocks"], cnt, ratio(gf[lv == 3], r[lv == 3]),
                                                                    ratio(gf, r)), flush=True)
        for w in sorted({W, 64}):
            Xf = X[:W // 2][:w] if w <= W // 2 else np.concatenate([X[:W // 2]] * 2)[:w]      # posterior-scale walkers only
            mc = timed(comp, Xf, names, const)
            t0 = time.perf_counter()
            for _ in range(20):
                comp.log_like_batch(Xf, names, const)
            print("W=%5d posterior-scale walkers, focused: prologue %.4f kernels %.4f epilogue %.4f ms; end to end %.4f ms; tables %s"
                  % (w, mc[0], mc[1], mc[2], (time.perf_counter() - t0) / 20 * 1e3, comp.engine.last_level_counts()), flush=True)
        comp.unfocus()
    for w in sorted({W, 64}):
        mc = timed(comp, X[:w], names, const)
        mf = timed(full, X[:w], names, const)
        print("W=%5d  compressed: prologue %.4f kernels %.4f epilogue %.4f ms   full grid: %.4f %.4f %.4f ms   kernel ratio %.1f"
              % (w, mc[0], mc[1], mc[2], mf[0], mf[1], mf[2], mf[1] / mc[1]), flush=True)
        t0 = time.perf_counter()
        for _ in range(20):
            comp.log_like_batch(X[:w], names, const)
        tc = (time.perf_counter() - t0) / 20
        t0 = time.perf_counter()
        for _ in range(20):
            full.log_like_batch(X[:w], names, const)
        tf = (time.perf_counter() - t0) / 20
        print("         end to end through log_like_batch (host in / host out): compressed %.4f ms, full grid %.4f ms" % (tc * 1e3, tf * 1e3))


if __name__ == "__main__":
    main()
