OUT=gpurun_out/r2v_stress.txt; rm -f $OUT
for frac in 0.08 0.16 0.25 0.4; do
  echo "== CMP_PRECISE_FRAC=$frac (fast window kernel on the plain-bin prefix)" | tee -a $OUT
  BAJES_B200_CMP_PRECISE_FRAC=$frac python scripts/prof_calls.py C2 4096 5 | tee -a $OUT
  BAJES_B200_CMP_PRECISE_FRAC=$frac python scripts/stress_tail.py --walkers 2000000 --modes mufu --config C2 --approx TaylorF2_3.5PN --boxes narrow --out $OUT --tag "C2 narrow frac=$frac" | grep -E "^C2 "
done
