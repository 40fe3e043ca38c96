python -m pytest tests -m gpu -x -q 2>&1 | tail -4
OUT=gpurun_out/r02_stress_compress.txt; rm -f $OUT
python scripts/stress_compress.py --walkers 2000000 --extreme 200000 --out $OUT 2>&1 | grep -v "^#"
python scripts/compress_check.py --config C2 --walkers 4096 2>&1 | tee gpurun_out/r02_compress_check_c2.txt | tail -12
python scripts/compress_check.py --config C2_55 --walkers 4096 2>&1 | tee gpurun_out/r02_compress_check_c255.txt | tail -5
