python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash scripts/gpu_run_traffic.sh > gpurun_out/r2g_traffic.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -3 gpurun_out/r02_bench_1gpu.err
python - <<'P'
import json
for line in open('gpurun_out/r02_bench_1gpu.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','parity_max_err_over_tol','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'])
        print('roofline',{k:d['roofline'][k] for k in ('bound','frac','kernel_ms','traffic')}, d['full_grid']['roofline']['frac'], d['focused']['ms_per_step'])
        for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('e2e',{}).get('value'), v.get('parity_max_err_over_tol'))
P
