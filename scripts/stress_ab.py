import os, sys, glob
import numpy as np
sys.path.insert(0, ".")
def tol(ref): return 1e-6 * np.abs(ref) + 1e-4
libs = sys.argv[1:]
import subprocess, json
if os.environ.get("STRESS_CHILD"):
    from bajes_b200 import GWLikelihood, synthetic as syn
    out = {}
    for name, W in (("C2_small", 100003), ("C1", 100003)):
        cfg = syn.CONFIGS[name]
        cl = syn.product_classes()
        net = syn.build_network(cfg, cl)
        Xa, names = syn.draw_walkers(cfg, W // 2, 11, "narrow")
        Xb, _ = syn.draw_walkers(cfg, W - W // 2, 12, "wide")
        X = np.concatenate([Xa, Xb]); const = syn.constants(cfg)
        b = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"], sincos="fp64").log_like_batch(X, names, const)
        for sc in ("mufu", "poly32"):
            a = GWLikelihood(*net, cfg["srate"], cfg["seglen"], cfg["approx"], sincos=sc).log_like_batch(X, names, const)
            fin = np.isfinite(b)
            r = np.abs(a[fin] - b[fin]) / tol(b[fin])
            out[name + ":" + sc] = [float(r.max()), float(np.sort(r)[-5]), float(np.quantile(r, 0.999)), float(np.sqrt(np.mean((a[fin]-b[fin])**2)))]
    print(json.dumps(out))
else:
    for lib in libs:
        env = dict(os.environ, STRESS_CHILD="1", BAJES_B200_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(os.path.basename(lib), line[0] if line else r.stderr[-400:])
