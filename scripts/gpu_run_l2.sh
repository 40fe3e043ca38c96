OUT=gpurun_out/r02_stress_compress_1e7.txt; rm -f $OUT
python scripts/stress_compress.py --walkers 10000000 --extreme 0 --out $OUT 2>&1 | grep -v "^#"
