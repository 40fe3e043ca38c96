python -m pytest tests/test_gpu_compress.py -x -q -s 2>&1 | tail -25
python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -3 gpurun_out/r2w_bench.err
python - <<'P'
import json
for line in open('gpurun_out/r2w_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','evals_per_s','parity_max_err_over_tol','gpu_launches')})
        print('e2e',d['e2e']); print('roofline',{k:d['roofline'][k] for k in ('bound','frac','kernel_ms','units_per_launch','instr_per_unit')})
        fg=d['full_grid']; print('full_grid',{k:fg[k] for k in ('value','ms_per_step','kernel_ms','compressed_vs_full_grid_max_err_over_tol')}, fg.get('e2e'), fg['roofline']['frac'])
        print('cpu',d['cpu_baseline'])
        for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('e2e',{}).get('value'), v.get('parity_max_err_over_tol'))
P
