set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_rot.json 2> gpurun_out/r2a_bench_rot.err; tail -c 600 gpurun_out/r2a_bench_rot.err
BAJES_B200_NO_ROT=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_norot.json 2> gpurun_out/r2a_bench_norot.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sincos poly32 > gpurun_out/r2a_bench_rot_poly.json 2>/dev/null
python - <<'PY'
import json
for n in ("rot","norot","rot_poly"):
    try:
        d=json.load(open("gpurun_out/r2a_bench_%s.json"%n))
        print(n, d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"])
    except Exception as e: print(n, "failed", e)
PY
python scripts/stress_tail.py --walkers 1000000 --modes mufu,poly32 --out gpurun_out/r2a_stress_rot.txt --tag rot
BAJES_B200_NO_ROT=1 python scripts/stress_tail.py --walkers 1000000 --modes mufu,poly32 --out gpurun_out/r2a_stress_norot.txt --tag norot
