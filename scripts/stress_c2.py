import sys, numpy as np
sys.path.insert(0, ".")
from bajes_b200 import GWLikelihood, synthetic as syn
def tol(ref): return 1e-6 * np.abs(ref) + 1e-4
cfg = syn.CONFIGS["C2"]
cl = syn.product_classes()
net = syn.build_network(cfg, cl)
W = 65537
Xa, names = syn.draw_walkers(cfg, W // 2, 21, "narrow")
Xb, _ = syn.draw_walkers(cfg, W - W // 2, 22, "wide")
X = np.concatenate([Xa, Xb]); const = syn.constants(cfg)
for approx in ("TaylorF2_3.5PN", "TaylorF2_5.5PN_7.5PNTides2020"):
    b = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos="fp64").log_like_batch(X, names, const)
    for sc in ("mufu", "poly32"):
        a = GWLikelihood(*net, cfg["srate"], cfg["seglen"], approx, sincos=sc).log_like_batch(X, names, const)
        fin = np.isfinite(b); r = np.abs(a[fin]-b[fin])/tol(b[fin])
        print("C2 full %-30s %-6s W=%d finite=%d ok=%s max=%.3f 5th=%.3f p99.9=%.3f" % (approx, sc, W, fin.sum(), np.array_equal(np.isfinite(a), fin), r.max(), np.sort(r)[-5], np.quantile(r, .999)), flush=True)
