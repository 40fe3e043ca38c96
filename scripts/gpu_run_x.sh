python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2x_bench_2gpu.json 2> gpurun_out/r2x_bench_2gpu.err; tail -3 gpurun_out/r2x_bench_2gpu.err
python - <<'P'
import json
for line in open('gpurun_out/r2x_bench_2gpu.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step']); print('strong',d['strong_scaling'])
        print('full',d['full_grid']['value'],d['full_grid']['ms_per_step'])
        for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('e2e',{}).get('value'), v.get('parity_max_err_over_tol'))
P
