#!/usr/bin/env python
"""
Where does the error of the worst walkers come from?  Finds the worst walkers of one stress batch (production kernel vs
the all-FP64 kernel), then evaluates them chunk by chunk of the frequency axis (bj_partial_device) and prints the error of
Re<d|h> per chunk and detector next to the chunk's own contribution.

  python scripts/diag_tail.py [--config C2_small] [--approx TaylorF2_3.5PN] [--mode mufu] [--walkers 262144] [--top 3]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bajes_b200 import GWLikelihood, synthetic as syn  # noqa: E402


def chunk_profile(like, pm, xd):
    """complex <d|h> contribution of every chunk: [n_chunks, W, n_ifo]"""
    dev = xd.device
    W, nifo = xd.shape[0], len(like.ifos)
    nch = like.engine.num_chunks()
    st = torch.cuda.current_stream(dev).cuda_stream
    z = torch.empty((nch, 2 * nifo, W), dtype=torch.float64, device=dev)
    for c in range(nch):
        like.engine.partial_device(xd.data_ptr(), W, xd.shape[1], pm, c, c + 1, z[c].data_ptr(), st)
    out = torch.empty(W, dtype=torch.float64, device=dev)
    parts = torch.empty((W, nifo, 3), dtype=torch.float64, device=dev)
    like.engine.loglike_device(xd.data_ptr(), W, xd.shape[1], pm, out.data_ptr(), parts.data_ptr(), st)
    torch.cuda.synchronize(dev)
    z = z.cpu().numpy()
    zc = z[:, 0::2, :] + 1j * z[:, 1::2, :]                 # [nch, nifo, W]
    tot = zc.sum(axis=0)                                     # [nifo, W]
    p = parts.cpu().numpy()
    fin = (p[:, :, 0] + 1j * p[:, :, 1]).T                   # [nifo, W]
    fac = fin / tot
    return (zc * fac[None]).transpose(0, 2, 1), out.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2_small")
    ap.add_argument("--approx", default="TaylorF2_3.5PN")
    ap.add_argument("--mode", default="mufu")
    ap.add_argument("--walkers", type=int, default=1 << 18)
    ap.add_argument("--top", type=int, default=3)
    ap.add_argument("--seed", type=int, default=1000)
    a = ap.parse_args()
    cfg = syn.CONFIGS[a.config]
    net = syn.build_network(cfg, syn.product_classes())
    const = syn.constants(cfg)
    dev = torch.device("cuda", 0)
    names = syn.WALKER_NAMES
    like = GWLikelihood(*net, cfg["srate"], cfg["seglen"], a.approx, sincos=a.mode)
    ref = GWLikelihood(*net, cfg["srate"], cfg["seglen"], a.approx, sincos="fp64")
    X, _ = syn.draw_walkers(cfg, a.walkers, a.seed, "narrow")
    got = like.log_like_batch(X, names, const)
    exact = ref.log_like_batch(X, names, const)
    r = np.abs(got - exact) / (1e-6 * np.abs(exact) + 1e-4)
    order = np.argsort(-r)[:a.top]
    typical = np.argsort(r)[len(r) // 2: len(r) // 2 + 1]
    sel = np.concatenate([order, typical])
    xd = torch.from_numpy(np.ascontiguousarray(X[sel])).to(dev)
    pa, la = chunk_profile(like, like.param_map(names, const), xd)
    pb, lb = chunk_profile(ref, ref.param_map(names, const), xd)
    nch = pa.shape[0]
    f = np.asarray(net[4])
    band = f[(f >= cfg["f_min"]) & (f <= cfg["f_max"])]
    per = int(np.ceil(len(band) / nch))
    print("# %s %s %s: %d chunks of ~%d bins; rows = chunk, columns per detector: error of Re<d|h> (x 1e-6) | chunk's Re<d|h>"
          % (a.config, a.approx, a.mode, nch, per))
    for j, w in enumerate(sel):
        d = (pa[:, j, :] - pb[:, j, :]).real
        print("walker %d  ratio %.3f  logL %.6g  dlogL %.3g  sum of chunk errors per det %s  x=%s"
              % (w, r[w], exact[w], got[w] - exact[w], np.array2string(d.sum(axis=0), precision=3),
                 np.array2string(X[w], precision=6, max_line_width=300)))
        cum = np.cumsum(d, axis=0)
        for c in range(0, nch, max(1, nch // 32)):
            c1 = min(nch, c + max(1, nch // 32))
            blk = d[c:c1].sum(axis=0)
            val = pb[c:c1, j, :].real.sum(axis=0)
            print("   chunks %3d-%3d  f~%7.1f Hz  err[1e-6] %s   Re<d|h> %s   cumulative err[1e-6] %s"
                  % (c, c1 - 1, band[min(len(band) - 1, c * per)], np.array2string(blk * 1e6, precision=2, suppress_small=True),
                     np.array2string(val, precision=1, suppress_small=True),
                     np.array2string(cum[c1 - 1] * 1e6, precision=2, suppress_small=True)))


if __name__ == "__main__":
    main()
