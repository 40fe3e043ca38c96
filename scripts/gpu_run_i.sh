python -m pytest tests -m gpu -x -q 2>&1 | tail -6
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err; tail -4 gpurun_out/r2i_bench_2gpu.err
