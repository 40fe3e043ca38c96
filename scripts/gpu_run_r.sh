python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r2r_bench_8gpu.json 2> gpurun_out/r2r_bench_8gpu.err; tail -4 gpurun_out/r2r_bench_8gpu.err
