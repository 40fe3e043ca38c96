set -x
python scripts/compress_check.py --config C2 --walkers 4096 2>&1 | tee gpurun_out/r2t_compress_c2.txt
python scripts/compress_check.py --config C2_small --walkers 4096 2>&1 | tee gpurun_out/r2t_compress_c2small.txt
python scripts/compress_check.py --config C1 --walkers 4096 2>&1 | tee gpurun_out/r2t_compress_c1.txt
python scripts/compress_check.py --config C2_55 --walkers 4096 2>&1 | tee gpurun_out/r2t_compress_c255.txt
OUT=gpurun_out/r2t_stress_compress.txt; rm -f $OUT
python scripts/stress_tail.py --walkers 1000000 --modes mufu --config C2 --approx TaylorF2_3.5PN --out $OUT --tag "compressed axis, C2 full size" | grep -E "^C2 |tol "
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
