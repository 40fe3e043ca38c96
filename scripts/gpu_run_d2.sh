OUT=gpurun_out/r2d2_nodes.txt; rm -f $OUT
for cfg in "32 10" "48 22" "64 32"; do
  set -- $cfg
  echo "== nodes per run $1, reach $2 rad" | tee -a $OUT
  BAJES_B200_COMPRESS_NODES=$1 BAJES_B200_COMPRESS_REACH=$2 python scripts/prof_calls.py C2 4096 5 | tee -a $OUT
  BAJES_B200_COMPRESS_NODES=$1 BAJES_B200_COMPRESS_REACH=$2 python scripts/stress_tail.py --walkers 2000000 --modes mufu --config C2 --approx TaylorF2_3.5PN --boxes narrow --out $OUT --tag "C2 narrow n=$1 X=$2" | grep -E "^C2 "
done
