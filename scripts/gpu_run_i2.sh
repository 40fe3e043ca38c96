python -m pytest tests/test_gpu_compress.py -x -q -k focus 2>&1 | tail -15
python scripts/compress_check.py --config C2 --walkers 4096 2>&1 | grep -i "focus\|level"
