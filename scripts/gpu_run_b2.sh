N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; tail -3 gpurun_out/r02_bench_${N}gpu.err
python - <<P
import json
for line in open('gpurun_out/r02_bench_${N}gpu.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step']); print('strong',d['strong_scaling'])
        print('full',d['full_grid']['value'],d['full_grid']['ms_per_step'])
        for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('e2e',{}).get('value'), v.get('parity_max_err_over_tol'), v.get('chain_bit_identical_to_one_gpu'))
        print(d['secondary']['C5']['ptmcmc']['value'], d['secondary']['C5']['ptmcmc']['chain_bit_identical_to_one_gpu'])
P
if [ "$2" = "test" ]; then python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3; fi
