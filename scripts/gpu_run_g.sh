python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err; tail -5 gpurun_out/r2g_bench_1gpu.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err; tail -8 gpurun_out/r2g_bench_2gpu.err
