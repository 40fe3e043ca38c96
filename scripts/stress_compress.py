#!/usr/bin/env python
"""
Compressed frequency axis under stress (C2 size: 128 s, 521 729 bins, H1/L1/V1), every walker against the all-FP64 kernel
on every bin:
  * an EXTREME box that exercises the routing: chirp mass 0.05 - 30 (log-uniform: sub-design, wide-phase and heavy
    systems), q to 8, spins to 0.9, tides to 5000, time shifts to +-1.5 s -- walkers land on level 0, level 1 and the full
    grid; worst |dlogL| / tol per table;
  * posterior- and prior-scale boxes, N walkers per approximant (all five).

  python scripts/stress_compress.py [--walkers 1000000] [--extreme 200000] [--out file]
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


def device_eval(like, pm, xd):
    out = torch.empty(xd.shape[0], dtype=torch.float64, device=xd.device)
    like.engine.loglike_device(xd.data_ptr(), xd.shape[0], xd.shape[1], pm, out.data_ptr(), 0,
                               torch.cuda.current_stream(xd.device).cuda_stream)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=1_000_000)
    ap.add_argument("--extreme", type=int, default=200_000)
    ap.add_argument("--batch", type=int, default=1 << 17)
    ap.add_argument("--approx", default=",".join(ALL))
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    fh = open(a.out, "a") if a.out else None

    def emit(s):
        print(s, flush=True)
        if fh:
            fh.write(s + "\n")
            fh.flush()

    cfg = syn.CONFIGS["C2"]
    net = syn.build_network(cfg, syn.product_classes())
    const, names = syn.constants(cfg), syn.WALKER_NAMES
    dev = torch.device("cuda", 0)
    emit("# stress_compress: C2 (521 729 bins), default context (compressed axes) vs sincos='fp64' (every bin, all FP64)")
    for approx in a.approx.split(","):
        like = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx)
        ref = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos="fp64")
        pm, pr = like.param_map(names, const), ref.param_map(names, const)
        info = [like.engine.compress_info(lv) for lv in range(2)]
        # ---- extreme box -------------------------------------------------------------------------------------------
        if a.extreme and approx in (ALL[0], ALL[-1]):
            W = a.extreme
            rng = np.random.default_rng(17)
            cols = dict(mchirp=np.exp(rng.uniform(np.log(0.05), np.log(30.0), W)), q=rng.uniform(1, 8, W),
                        s1z=rng.uniform(-0.9, 0.9, W), s2z=rng.uniform(-0.9, 0.9, W), lambda1=rng.uniform(0, 5000, W),
                        lambda2=rng.uniform(0, 5000, W), distance=np.exp(rng.uniform(0, np.log(1000), W)),
                        cos_iota=rng.uniform(-1, 1, W), ra=rng.uniform(0, 2 * np.pi, W), dec=np.arcsin(rng.uniform(-1, 1, W)),
                        psi=rng.uniform(0, np.pi, W), time_shift=rng.uniform(-1.5, 1.5, W), phi_ref=rng.uniform(0, 2 * np.pi, W))
            X = np.stack([cols[k] for k in names], axis=1)
            worst = np.zeros(3)
            counts = np.zeros(3, dtype=np.int64)
            ok = True
            for i0 in range(0, W, a.batch):
                xd = torch.from_numpy(X[i0:i0 + a.batch]).to(dev)
                got, want = device_eval(like, pm, xd), device_eval(ref, pr, xd)
                torch.cuda.synchronize(dev)
                level = torch.as_tensor(like.engine.last_levels(xd.shape[0]), device=dev)
                fin = torch.isfinite(want)
                ok = ok and bool(torch.equal(torch.isfinite(got), fin))
                r = torch.where(fin, (got - want).abs() / (1e-6 * want.abs() + 1e-4), torch.zeros_like(want))
                for lv in range(3):
                    sel = level == (lv if lv < 2 else 2)
                    counts[lv] += int(sel.sum())
                    if bool(sel.any()):
                        worst[lv] = max(worst[lv], float(r[sel].max()))
            emit("%-34s EXTREME box n=%d pattern_ok=%s  walkers (level 0 / level 1 / full grid) %s  worst |d|/tol per table "
                 "%.3f / %.3f / %.3f" % (approx, W, ok, counts.tolist(), worst[0], worst[1], worst[2]))
        # ---- posterior / prior boxes --------------------------------------------------------------------------------
        done, ib, t0, t_c, t_r = 0, 0, time.time(), 0.0, 0.0
        mx, n_fin, ok, lv_counts = [], 0, True, np.zeros(3, dtype=np.int64)
        hist = np.zeros(41, dtype=np.int64)
        while done < a.walkers:
            n = min(a.batch, a.walkers - done)
            X, _ = syn.draw_walkers(cfg, n, 7000 + ib, ("narrow", "wide")[ib % 2])
            ib += 1
            xd = torch.from_numpy(X).to(dev)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            got = device_eval(like, pm, xd)
            e[1].record()
            want = device_eval(ref, pr, xd)
            e[2].record()
            torch.cuda.synchronize(dev)
            t_c += e[0].elapsed_time(e[1])
            t_r += e[1].elapsed_time(e[2])
            lv_counts += np.asarray(like.engine.last_level_counts())
            fin = torch.isfinite(want)
            ok = ok and bool(torch.equal(torch.isfinite(got), fin))
            r = torch.where(fin, (got - want).abs() / (1e-6 * want.abs() + 1e-4), torch.zeros_like(want))
            n_fin += int(fin.sum())
            hist += torch.histc(r.clamp(max=2.0).float(), bins=41, min=0.0, max=2.05).long().cpu().numpy()
            mx += torch.topk(r, 5).values.tolist()
            mx = sorted(mx, reverse=True)[:5]
            done += n
        cum = np.cumsum(hist) / max(1, hist.sum())
        q = lambda p: 0.05 * (int(np.searchsorted(cum, p)) + 1)
        emit("%-34s n=%d finite=%d pattern_ok=%s nodes %d/%d  tables %s  max|d|/tol=%.3f 5th=%.3f p99.9<=%.2f p99.999<=%.2f  "
             "(%.0f ms compressed, %.0f ms fp64 full grid, %.1f s)"
             % (approx, done, n_fin, ok, info[0]["nodes"], info[1]["nodes"], lv_counts.tolist(), mx[0], mx[-1], q(0.999),
                q(0.99999), t_c, t_r, time.time() - t0))
        like.engine.close()
        ref.engine.close()


if __name__ == "__main__":
    main()
