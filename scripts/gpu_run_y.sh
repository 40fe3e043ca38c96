python -m pytest tests/test_gpu_compress.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -6
python scripts/compress_check.py --config C2 --walkers 4096 2>&1 | tee gpurun_out/r2y_compress_c2.txt | tail -9
python scripts/compress_check.py --config C2_55 --walkers 4096 2>&1 | tee gpurun_out/r2y_compress_c255.txt | tail -7
