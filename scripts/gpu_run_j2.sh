python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -3 gpurun_out/r02_bench_1gpu.err
python - <<'P'
import json
for line in open('gpurun_out/r02_bench_1gpu.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','parity_max_err_over_tol','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'])
        print('roofline',{k:d['roofline'][k] for k in ('bound','frac','kernel_ms','traffic')})
        print('focused', json.dumps(d['focused'])[:900])
        print('full_grid', d['full_grid']['ms_per_step'], d['full_grid']['roofline']['frac'], d['full_grid']['roofline']['traffic'])
P
python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1
