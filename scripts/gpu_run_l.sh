python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for ch in 128 256 512; do
echo "== chunks $ch"
BAJES_B200_CHUNKS=$ch python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from bajes_b200 import GWLikelihood, synthetic as syn
for name in ("C2",):
    cfg = syn.CONFIGS[name]
    like = GWLikelihood(*syn.build_network(cfg, syn.product_classes()), cfg["srate"], cfg["seglen"], cfg["approx"])
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    for W in (64, 192, 512, 4096):
        X, _ = syn.draw_walkers(cfg, W, 5, "narrow")
        for _ in range(5): like.log_like_batch(X, names, const)
        t = []
        for _ in range(100):
            t0 = time.perf_counter(); like.log_like_batch(X, names, const); t.append(time.perf_counter() - t0)
        print(name, "W=%d" % W, "host API median %.4f ms min %.4f ms" % (1e3*np.median(t), 1e3*np.min(t)), "kernel ms", [round(x,4) for x in like.engine.last_kernel_ms()], like.engine.last_launch_info(), flush=True)
PY
done
