import sys, numpy as np
sys.path.insert(0, '.')
from bajes_b200 import GWLikelihood, synthetic as syn
cfg = syn.CONFIGS["C2"]
cl = syn.product_classes()
like = GWLikelihood(*syn.build_network(cfg, cl), cfg["srate"], cfg["seglen"], cfg["approx"], marg_time_shift=True)
X, names = syn.draw_walkers(cfg, 512, 3, "narrow")
const = syn.constants(cfg)
for _ in range(2):
    like.log_like_batch(X, names, const)
print(like.engine.last_kernel_ms(), like.engine.last_launch_info())
