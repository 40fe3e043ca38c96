python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -3 gpurun_out/r2z_bench.err
python - <<'P'
import json
for line in open('gpurun_out/r2z_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','evals_per_s','parity_max_err_over_tol','gpu_launches')})
        print('e2e',d['e2e']['value'],d['e2e']['ms_per_step']); print('roofline',{k:d['roofline'][k] for k in ('bound','frac','kernel_ms','units_per_launch','instr_per_unit')})
        fg=d['full_grid']; print('full_grid',{k:fg[k] for k in ('value','ms_per_step','kernel_ms','compressed_vs_full_grid_max_err_over_tol')}, fg['roofline']['frac'])
        for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('e2e',{}).get('value'), v.get('parity_max_err_over_tol'))
P
python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_c2_4096.csv python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1
python scripts/prof_calls.py C2 64 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_c2_64.csv python scripts/prof_calls.py C2 64 3 > /dev/null 2>&1
