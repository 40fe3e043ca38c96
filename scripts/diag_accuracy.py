"""GPU diagnostic: accuracy of each sin/cos mode against the reference golden vectors + first timings.
Usage (GPU box): python scripts/diag_accuracy.py [--full]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bajes_b200 import GWLikelihood, engine, synthetic as syn   # noqa: E402
from conftest import load_golden, port_network, product_network_from_port   # noqa: E402
from oracle import harness   # noqa: E402


def run(cfg_name, gold_name, key, full_timing=False):
    cfg = syn.CONFIGS[cfg_name]
    gold = load_golden(gold_name)
    net = port_network(cfg, gold)
    pnet = product_network_from_port(cfg, net)
    names = [str(n) for n in gold["names"]]
    const = syn.constants(cfg)
    ref = gold[key]
    X = gold["X"][:len(ref)]
    fin = np.isfinite(ref)
    tol = harness.tolerance(ref[fin])
    for mode in ("fp64", "poly32", "mufu"):
        like = GWLikelihood(*pnet, cfg["srate"], cfg["seglen"], cfg["approx"], sincos=mode)
        got = like.log_like_batch(X, names, const)
        err = (got[fin] - ref[fin])
        r = np.abs(err) / tol
        print("%-9s %-7s n=%d  max|d|/tol=%.3g  median=%.3g  mean(d)=%.3g  rms(d)=%.3g  max|d|=%.3g"
              % (cfg_name, mode, fin.sum(), r.max(), np.median(r), err.mean(), np.sqrt((err ** 2).mean()),
                 np.abs(err).max()), flush=True)
        if full_timing:
            Xb, _ = syn.draw_walkers(cfg, cfg["n_walkers"], 5, "narrow")
            like.log_like_batch(Xb, names, const)
            t0 = time.time()
            like.log_like_batch(Xb, names, const)
            wall = time.time() - t0
            ms = like.engine.last_kernel_ms()
            nb = like.engine.n_bins
            print("    timing W=%d N=%d: prologue %.3f ms, main %.3f ms, epilogue %.3f ms, wall %.3f ms -> %.3g walker*bins/s (kernel), launch %s"
                  % (len(Xb), nb, ms[0], ms[1], ms[2], wall * 1e3, len(Xb) * nb / (ms[1] * 1e-3),
                     like.engine.last_launch_info()), flush=True)
        like.engine.close()


if __name__ == "__main__":
    print(engine.device_info(0))
    print("peak DFMA thread-instr/s: %.4g   MUFU: %.4g" % (engine.measure_peak("dfma"), engine.measure_peak("mufu")))
    run("C1", "c1_bbh.npz", "logl_marg0")
    run("C2_small", "c2_small.npz", "logl_TaylorF2_3.5PN_marg0")
    if "--full" in sys.argv:
        run("C2", "c2.npz", "logl_TaylorF2_3.5PN_marg0", full_timing=True)
