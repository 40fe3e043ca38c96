#!/usr/bin/env python
"""
Extreme-value stress of the production kernels: N walkers per case (default 10^7, in batches), every walker compared
with the all-FP64 validation kernel of the same library on the same GPU.  Reports the worst |dlogL| / tol, the 5th worst,
quantiles, and the inner products of the worst walkers (which detector / which quantity carries the error).

  python scripts/stress_tail.py [--walkers 10000000] [--batch 262144] [--config C2_small] [--approx a,b,...]
                                [--modes mufu,poly32] [--out profiles/rXX_stress_tail.txt]

Tolerance: |dlogL| <= 1e-6 |logL| + 1e-4 (BASELINE.json north star).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bajes_b200 import GWLikelihood, synthetic as syn  # noqa: E402

ALL = ["TaylorF2_3.5PN", "TaylorF2_5.5PN", "TaylorF2_5.5PN_7.5PNTides", "TaylorF2_5.5PN_3.5PNQM_7.5PNTides",
       "TaylorF2_5.5PN_7.5PNTides2020"]


def tol(ref):
    return 1e-6 * np.abs(ref) + 1e-4


def device_eval(like, pm, xd, want_parts):
    dev = xd.device
    out = torch.empty(xd.shape[0], dtype=torch.float64, device=dev)
    parts = torch.empty((xd.shape[0], len(like.ifos), 3), dtype=torch.float64, device=dev) if want_parts else None
    like.engine.loglike_device(xd.data_ptr(), xd.shape[0], xd.shape[1], pm, out.data_ptr(),
                               parts.data_ptr() if want_parts else 0, torch.cuda.current_stream(dev).cuda_stream)
    return out, parts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=1 << 18)
    ap.add_argument("--config", default="C2_small")
    ap.add_argument("--approx", default=",".join(ALL))
    ap.add_argument("--modes", default="mufu")
    ap.add_argument("--boxes", default="narrow,wide")
    ap.add_argument("--marg-phi", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    fh = open(a.out, "a") if a.out else None

    def emit(s):
        print(s, flush=True)
        if fh:
            fh.write(s + "\n")
            fh.flush()

    cfg = syn.CONFIGS[a.config]
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    const = syn.constants(cfg)
    dev = torch.device("cuda", 0)
    boxes = a.boxes.split(",")
    emit("# stress_tail %s: %s, %d walkers per case in batches of %d, boxes %s, marg_phi=%s, env ROT=%s"
         % (a.tag, a.config, a.walkers, a.batch, boxes, a.marg_phi, "off" if os.environ.get("BAJES_B200_NO_ROT") else "on"))
    kw = dict(marg_phi_ref=True) if a.marg_phi else {}
    for approx in a.approx.split(","):
        ref_like = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos="fp64", **kw)
        for mode in a.modes.split(","):
            like = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos=mode, **kw)
            names = syn.WALKER_NAMES
            pm, pr = like.param_map(names, const), ref_like.param_map(names, const)
            done, t0 = 0, time.time()
            worst = []                      # (ratio, x, logl_ref, d, parts_mode, parts_ref)
            hist = np.zeros(41, dtype=np.int64)       # ratio histogram, bins of 0.05
            pattern_ok, n_fin, t_mode, t_ref, ib = True, 0, 0.0, 0.0, 0
            while done < a.walkers:
                n = min(a.batch, a.walkers - done)
                box = boxes[ib % len(boxes)]
                X, _ = syn.draw_walkers(cfg, n, 1000 + ib, box)
                ib += 1
                xd = torch.from_numpy(X).to(dev)
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                got, gp = device_eval(like, pm, xd, True)
                e[1].record()
                ref, rp = device_eval(ref_like, pr, xd, True)
                e[2].record()
                torch.cuda.synchronize(dev)
                t_mode += e[0].elapsed_time(e[1])
                t_ref += e[1].elapsed_time(e[2])
                fin = torch.isfinite(ref)
                pattern_ok = pattern_ok and bool(torch.equal(torch.isfinite(got), fin))
                r = torch.where(fin, (got - ref).abs() / (1e-6 * ref.abs() + 1e-4), torch.zeros_like(ref))
                n_fin += int(fin.sum())
                hist += torch.histc(r.clamp(max=2.0).float(), bins=41, min=0.0, max=2.05).long().cpu().numpy()
                top = torch.topk(r, min(5, n))
                for v, j in zip(top.values.tolist(), top.indices.tolist()):
                    worst.append((v, X[j].copy(), float(ref[j]), float(got[j] - ref[j]), gp[j].cpu().numpy(),
                                  rp[j].cpu().numpy(), box))
                worst = sorted(worst, key=lambda w: -w[0])[:5]
                done += n
            cum = np.cumsum(hist) / max(1, hist.sum())
            q = lambda p: 0.05 * (int(np.searchsorted(cum, p)) + 1)
            emit("%-9s %-34s %-6s n=%d finite=%d pattern_ok=%s max|d|/tol=%.3f 5th=%.3f p99.9<=%.2f p99.999<=%.2f "
                 "(%.1f s: %.0f ms %s, %.0f ms fp64)"
                 % (a.config, approx, mode, done, n_fin, pattern_ok, worst[0][0], worst[-1][0], q(0.999), q(0.99999),
                    time.time() - t0, t_mode, mode, t_ref))
            for v, x, lr, d, gp, rp, box in worst[:3]:
                dd = gp - rp
                emit("    %.3f tol  box=%s logL=%.6g dlogL=%.3g  <h|h>=%s  d Re<d|h> per det=%s  rel=%s  x=%s"
                     % (v, box, lr, d, np.array2string(rp[:, 2], precision=1), np.array2string(dd[:, 0], precision=2),
                        np.array2string(dd[:, 0] / np.maximum(np.abs(rp[:, 0]), 1e-30), precision=2),
                        np.array2string(x, precision=7, max_line_width=400)))
            like.engine.close()
        ref_like.engine.close()


if __name__ == "__main__":
    main()
