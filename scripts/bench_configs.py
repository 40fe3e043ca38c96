#!/usr/bin/env python
"""
Secondary measurements (one GPU): the BASELINE.json configs other than the headline C2, plus kernel variants.
Writes one JSON object per line to stdout (and profiles/ when --out is given).

  C1    TaylorF2 3.5PN BBH, H1+L1, 4 s, 64 walkers                (launch-latency bound)
  C2    headline, per sin/cos mode and for the 5.5PN + 7.5PN-tides phase model
  C3    ROQ, 65 536 points, N_L = 400 / N_Q = 200 synthetic JenpyROQ-format basis, 10 000 tc points
  C4    relative binning, 256 s, 2^20 points (device-resident and host-resident)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bajes_b200 import GWLikelihood, synthetic as syn  # noqa: E402
from bajes_b200.binning import GWBinningLikelihood  # noqa: E402
from bajes_b200.ensemble import BoxPrior  # noqa: E402
from bajes_b200 import roq as roqmod  # noqa: E402

DEV = torch.device("cuda", 0)


def time_device(like, X, names, const, reps=5):
    pm = like.param_map(names, const)
    xd = torch.from_numpy(np.ascontiguousarray(X)).to(DEV)
    out = torch.empty(len(X), dtype=torch.float64, device=DEV)
    st = torch.cuda.current_stream(DEV).cuda_stream
    ms = []
    for i in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        like.engine.loglike_device(xd.data_ptr(), len(X), X.shape[1], pm, out.data_ptr(), 0, st)
        b.record()
        b.synchronize()
        if i >= 2:
            ms.append(a.elapsed_time(b))
    return float(np.median(ms)), like.engine.last_kernel_ms(), out.cpu().numpy()


def time_host(like, X, names, const, reps=3):
    like.log_like_batch(X, names, const)
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        like.log_like_batch(X, names, const)
        t.append(time.perf_counter() - t0)
    return float(np.median(t)) * 1e3


def emit(rec, fh):
    line = json.dumps(rec)
    print(line, flush=True)
    if fh:
        fh.write(line + "\n")
        fh.flush()


def fullgrid(name, fh, sincos="mufu", walkers=None, approx=None, nspcal=0, nweights=0, tmarg=False):
    cfg = dict(syn.CONFIGS[name])
    if approx:
        cfg["approx"] = approx
    cl = syn.product_classes()
    ex = syn.extras_setup(cfg, nspcal, nweights)
    like = GWLikelihood(*syn.build_network(cfg, cl), cfg["srate"], cfg["seglen"], cfg["approx"], sincos=sincos,
                        marg_time_shift=tmarg, **ex)
    W = walkers or cfg["n_walkers"]
    X, names = syn.draw_walkers(cfg, W, 5, "narrow")
    if nspcal or nweights:
        Xe, ne = syn.extras_columns(cfg, W, 6, nspcal, nweights)
        X, names = np.ascontiguousarray(np.concatenate([X, Xe], axis=1)), names + ne
    const = syn.constants(cfg)
    step_ms, kms, _ = time_device(like, X, names, const)
    host_ms = time_host(like, X, names, const)
    nb = like.engine.n_bins
    emit({"config": name, "kind": "fullgrid", "approx": cfg["approx"], "sincos": sincos, "walkers": W, "n_bins": nb,
          "nspcal": nspcal, "nweights": nweights, "ndim": X.shape[1], "marg_time_shift": tmarg,
          "n_ifo": len(cfg["ifos"]), "step_ms_device": step_ms, "kernel_ms": kms, "step_ms_host_api": host_ms,
          "walker_bins_per_s": W * nb / (step_ms * 1e-3), "evals_per_s": W / (step_ms * 1e-3),
          "launch": like.engine.last_launch_info()}, fh)
    like.engine.close()


def roq_c3(fh, points=65536, n_lin=400, n_quad=200, tc_points=10000, seglen=8.0):
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=seglen, seed=3003)
    cl = syn.product_classes()
    net = syn.build_network(cfg, cl)
    ifos, datas, dets, noises, f = net
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    basis = syn.make_roq_basis(cfg, n_lin=n_lin, n_quad=n_quad, seed=7)
    prior = BoxPrior(names, [[-1, 1]] * len(names), {})
    prior.bounds[names.index("time_shift")] = [-0.05, 0.05]
    import tempfile
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        syn.write_jenpyroq(tmp, basis, cfg)
        roq = roqmod.build_roq_dict(tmp, tc_points, ifos, datas, noises, f, cfg["f_min"], cfg["f_max"], cfg["seglen"],
                                    prior, cfg["approx"])
    setup_s = time.perf_counter() - t0
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"], roq=roq)
    nl, nq = len(roq["mask_omega"]), len(roq["mask_psi"])
    out = {}
    for box in ("narrow", "wide"):
        X, _ = syn.draw_walkers(cfg, points, 9, box)
        step_ms, kms, logl = time_device(like, X, names, const)
        host_ms = time_host(like, X, names, const)
        alg_bytes = points * len(ifos) * 4 * nl * 16
        out[box] = {"step_ms_device": step_ms, "kernel_ms": kms, "step_ms_host_api": host_ms,
                    "points_per_s": points / (step_ms * 1e-3), "algorithmic_GBps": alg_bytes / (kms[1] * 1e-3) / 1e9,
                    "finite_fraction": float(np.mean(np.isfinite(logl)))}
    emit({"config": "C3", "kind": "roq", "points": points, "n_lin": nl, "n_quad": nq, "tc_points": tc_points,
          "n_ifo": len(ifos), "n_bins_full": int(len(np.asarray(f)[datas[ifos[0]].mask])), "setup_s": setup_s,
          "spline_table_MB": tc_points * nl * 16 * len(ifos) / 1e6, "algorithmic_bytes_per_point": len(ifos) * 4 * nl * 16,
          **out}, fh)
    like.engine.close()


def relbin_c4(fh, name="C4", points=1 << 20):
    cfg = syn.CONFIGS[name]
    cl = syn.product_classes()
    ifos, datas, dets, noises, f = syn.build_network(cfg, cl)
    fid = syn.injection_params(cfg)
    t0 = time.perf_counter()
    like = GWBinningLikelihood(ifos, datas, dets, noises, dict(fid), f, cfg["srate"], cfg["seglen"], cfg["approx"])
    setup_s = time.perf_counter() - t0
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    X, _ = syn.draw_walkers(cfg, points, 11, "narrow")
    step_ms, kms, logl = time_device(like, X, names, const)
    host_ms = time_host(like, X, names, const)
    emit({"config": name, "kind": "relbin", "points": points, "n_bins_rb": int(like.Nbin), "seglen": cfg["seglen"],
          "n_ifo": len(ifos), "setup_s": setup_s, "step_ms_device": step_ms, "kernel_ms": kms,
          "step_ms_host_api": host_ms, "points_per_s_device": points / (step_ms * 1e-3),
          "points_per_s_host_api": points / (host_ms * 1e-3), "h2d_bytes": int(X.nbytes), "d2h_bytes": points * 8,
          "finite_fraction": float(np.mean(np.isfinite(logl)))}, fh)
    like.engine.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    fh = open(a.out, "w") if a.out else None
    only = set(a.only.split(",")) if a.only else None
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    want = lambda k: only is None or k in only
    if want("c1"):
        fullgrid("C1", fh)
        fullgrid("C1", fh, walkers=4096)
    if want("c2"):
        for mode in ("mufu", "poly32", "fp64"):
            fullgrid("C2", fh, sincos=mode)
        fullgrid("C2", fh, approx="TaylorF2_5.5PN_7.5PNTides")
        fullgrid("C2", fh, walkers=64)
        fullgrid("C2", fh, walkers=512)
    if want("extras"):
        fullgrid("C2", fh, nspcal=10)
        fullgrid("C2", fh, nweights=8)
        fullgrid("C2", fh, nspcal=10, nweights=8)
    if want("tmarg"):
        fullgrid("C2", fh, tmarg=True, walkers=1024)
        fullgrid("C1", fh, tmarg=True, walkers=4096)
    if want("c3"):
        roq_c3(fh)
    if want("c4"):
        relbin_c4(fh)
