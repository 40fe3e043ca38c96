python -m pytest tests -m gpu -x -q 2>&1 | tail -4
OUT=gpurun_out/r02_stress_compress.txt; rm -f $OUT
python scripts/stress_compress.py --walkers 1000000 --extreme 200000 --out $OUT 2>&1 | grep -v "^#" 
python scripts/compress_check.py --config C2 --walkers 4096 2>&1 | tee gpurun_out/r02_compress_check_c2.txt | tail -5
python bench.py --steps 10 --warmup 3 --secondary C5 --no-cpu-baseline > gpurun_out/r2a2_bench.json 2> gpurun_out/r2a2_bench.err; tail -3 gpurun_out/r2a2_bench.err
python - <<'P'
import json
for line in open('gpurun_out/r2a2_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step')}); c5=d['secondary']['C5']; print(c5['value'], c5['s_per_sampler_step']); print(json.dumps(c5['ptmcmc'],indent=1))
P
