python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --secondary C1 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from bajes_b200 import GWLikelihood, synthetic as syn
for name in ("C1", "C2"):
    cfg = syn.CONFIGS[name]
    like = GWLikelihood(*syn.build_network(cfg, syn.product_classes()), cfg["srate"], cfg["seglen"], cfg["approx"])
    names, const = syn.WALKER_NAMES, syn.constants(cfg)
    for W in (1, 64, 512):
        X, _ = syn.draw_walkers(cfg, W, 5, "narrow")
        for _ in range(5): like.log_like_batch(X, names, const)
        t = []
        for _ in range(200):
            t0 = time.perf_counter(); like.log_like_batch(X, names, const); t.append(time.perf_counter() - t0)
        print(name, "W=%d" % W, "host API median %.4f ms min %.4f ms" % (1e3*np.median(t), 1e3*np.min(t)), "kernel ms", [round(x,4) for x in like.engine.last_kernel_ms()], flush=True)
PY
