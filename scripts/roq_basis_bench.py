#!/usr/bin/env python
"""ROQ basis construction on the GPU at a serious size: 8 s segment, 20-2048 Hz (16 225 bins), BNS box, two stages of
training waveforms; reports basis sizes, greedy time and the HBM traffic it stands for (3 x matrix per basis vector)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bajes_b200 import synthetic as syn                     # noqa: E402
from bajes_b200.roq_basis import build_roq_basis            # noqa: E402

n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = dict(syn.CONFIGS["C2_small"])
cfg.update(seglen=8.0, srate=4096.0, f_min=20.0, f_max=2048.0)
f = syn.frequency_axis(cfg)
fb = f[(f >= cfg["f_min"]) & (f <= cfg["f_max"])]
fa, asd = syn.product_classes().design_asd("H1")
psd = np.interp(fb, fa, asd) ** 2
names = ["mchirp", "q", "s1z", "s2z", "lambda1", "lambda2", "distance", "cos_iota", "phi_ref"]
ranges = {"mchirp": (1.195, 1.200), "q": (1.0, 1.6), "s1z": (-0.05, 0.05), "s2z": (-0.05, 0.05), "lambda1": (0.0, 1500.0),
          "lambda2": (0.0, 1500.0), "distance": 100.0, "cos_iota": 1.0, "phi_ref": 0.0}
t0 = time.time()
b = build_roq_basis(fb, psd, cfg["seglen"], cfg["approx"], ranges, names, stages=(n_train, n_train), tol=1e-6, max_linear=768,
                    max_quadratic=256, seed=1)
wall = time.time() - t0
out = {"bins": len(fb), "training_per_stage": n_train, "wall_s": wall}
for tag in ("lin", "quad"):
    info = b[tag + "_info"]
    steps = sum(s["basis_after"] - s["basis_before"] + s["basis_before"] for s in info["stages"])     # projections run
    bytes_ = sum(3.0 * s["training"] * len(fb) * 16 * (s["basis_after"]) for s in info["stages"])
    out[tag] = {"nodes": int(b[tag + "_b"].shape[1]), "stages": info["stages"], "greedy_ms": info["greedy_ms"],
                "algorithmic_GBps": bytes_ / (info["greedy_ms"] * 1e-3) / 1e9}
print(json.dumps(out, indent=1))
