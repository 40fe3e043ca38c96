bash scripts/gpu_run_traffic.sh > gpurun_out/r2g_traffic.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -3 gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
python scripts/roq_basis_bench.py 8192 > gpurun_out/r02_roq_basis_bench.json 2>/dev/null
python - <<'P'
import json
for line in open('gpurun_out/r02_bench_1gpu.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','evals_per_s','parity_max_err_over_tol','gpu_launches')})
        print('e2e',d['e2e']['value'],d['e2e']['ms_per_step']); print('roofline',{k:d['roofline'][k] for k in ('bound','frac','kernel_ms','units_per_launch','instr_per_unit','traffic')})
        fg=d['full_grid']; print('full_grid',{k:fg[k] for k in ('value','ms_per_step','kernel_ms','compressed_vs_full_grid_max_err_over_tol')}, fg['roofline']['frac'], fg['roofline']['traffic'])
        print('cpu', d['cpu_baseline'])
        for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('e2e',{}).get('value'), v.get('parity_max_err_over_tol'))
P
