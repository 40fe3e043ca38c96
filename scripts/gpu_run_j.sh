python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
for n in 8 4; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 ) > gpurun_out/r2j_bench_${n}gpu.json 2> gpurun_out/r2j_bench_${n}gpu.err; tail -4 gpurun_out/r2j_bench_${n}gpu.err
done
