#!/usr/bin/env python
"""A few batched likelihood calls for the profiler: python scripts/prof_calls.py CONFIG WALKERS [CALLS] [KIND]
KIND: fullgrid (default) | roq | relbin"""
import sys

import numpy as np

sys.path.insert(0, ".")
from bajes_b200 import GWLikelihood, synthetic as syn   # noqa: E402

name, W = sys.argv[1], int(sys.argv[2])
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 3
kind = sys.argv[4] if len(sys.argv) > 4 else "fullgrid"
names = syn.WALKER_NAMES
if kind == "fullgrid":
    cfg = syn.CONFIGS[name]
    like = GWLikelihood(*syn.build_network(cfg, syn.product_classes()), cfg["srate"], cfg["seglen"], cfg["approx"])
    X, _ = syn.draw_walkers(cfg, W, 5, "narrow")
elif kind == "roq":
    import tempfile
    from bajes_b200 import roq as roqmod
    from bajes_b200.ensemble import BoxPrior
    cfg = dict(syn.CONFIGS["C2_small"])
    cfg.update(seglen=8.0, seed=3003)
    ifos, datas, dets, noises, f = syn.build_network(cfg, syn.product_classes())
    basis = syn.make_roq_basis(cfg, n_lin=400, n_quad=200, seed=7)
    prior = BoxPrior(names, [[-1, 1]] * len(names), {})
    prior.bounds[names.index("time_shift")] = [-0.05, 0.05]
    with tempfile.TemporaryDirectory() as tmp:
        syn.write_jenpyroq(tmp, basis, cfg)
        roq = roqmod.build_roq_dict(tmp, 10000, ifos, datas, noises, f, cfg["f_min"], cfg["f_max"], cfg["seglen"], prior,
                                    cfg["approx"], device="cuda:0")
    like = GWLikelihood(ifos, datas, dets, noises, f, cfg["srate"], cfg["seglen"], cfg["approx"], roq=roq)
    X, _ = syn.draw_walkers(cfg, W, 9, "wide")
else:
    from bajes_b200.binning import GWBinningLikelihood
    cfg = syn.CONFIGS["C4"]
    ifos, datas, dets, noises, f = syn.build_network(cfg, syn.product_classes())
    like = GWBinningLikelihood(ifos, datas, dets, noises, dict(syn.injection_params(cfg)), f, cfg["srate"], cfg["seglen"],
                               cfg["approx"])
    X, _ = syn.draw_walkers(cfg, W, 11, "narrow")
const = syn.constants(cfg)
for _ in range(calls):
    out = like.log_like_batch(X, names, const)
print(kind, name, W, "calls", calls, "logL[0]", out[0], "kernel ms", like.engine.last_kernel_ms())
