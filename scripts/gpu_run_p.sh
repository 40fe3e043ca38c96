python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "roq or relbin" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --secondary C3,C4 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -3 gpurun_out/r2p_bench.err
