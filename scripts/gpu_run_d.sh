python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for fr in 0 0.03 0.06; do
  BAJES_B200_PRECISE_FRAC=$fr python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_$fr.json 2>/dev/null
done
for fr in 0.03 0.06; do
  BAJES_B200_PRECISE_FRAC=$fr python scripts/stress_tail.py --walkers 1000000 --modes mufu --approx TaylorF2_3.5PN,TaylorF2_5.5PN_7.5PNTides2020 --out gpurun_out/r2d_stress_$fr.txt --tag frac$fr | grep -E "^C2_small"
done
python - <<'PY'
import json
for n in ("0","0.03","0.06"):
    d=json.load(open("gpurun_out/r2d_bench_%s.json"%n))
    print(n, d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"])
PY
