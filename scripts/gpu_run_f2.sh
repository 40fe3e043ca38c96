OUT=gpurun_out/r2f2_window.txt; rm -f $OUT
for frac in 0.16 0.22 0.27 0.35; do
  echo "== CMP_PRECISE_FRAC=$frac (13312 nodes)" | tee -a $OUT
  BAJES_B200_CMP_PRECISE_FRAC=$frac python scripts/prof_calls.py C2 4096 5 | tee -a $OUT
  BAJES_B200_CMP_PRECISE_FRAC=$frac python scripts/stress_tail.py --walkers 4000000 --modes mufu --config C2 --approx TaylorF2_3.5PN --boxes narrow --out $OUT --tag "C2 narrow frac=$frac" | grep -E "^C2 "
done
