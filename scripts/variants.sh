for lib in libbajes_b200.so libbajes_b200_t384_2.so libbajes_b200_t512_2.so libbajes_b200_t320_3.so; do
  BAJES_B200_LIB=$PWD/bajes_b200/csrc/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib kernel_ms=%.3f'%d['roofline']['kernel_ms'])"
done
