python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or extras or launch" 2>&1 | tail -3
for sh in 0.8 0.9 0.95; do
  BAJES_B200_PRECISE_SHARE=$sh python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_$sh.json 2>/dev/null
  BAJES_B200_PRECISE_SHARE=$sh python scripts/stress_tail.py --walkers 1000000 --modes mufu --approx TaylorF2_3.5PN,TaylorF2_5.5PN_7.5PNTides2020 --out gpurun_out/r2e_stress_$sh.txt --tag share$sh | grep -E "^C2_small"
done
python - <<'PY'
import json
for n in ("0.8","0.9","0.95"):
    d=json.load(open("gpurun_out/r2e_bench_%s.json"%n))
    print(n, d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"])
PY
