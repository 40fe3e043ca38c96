mkdir -p /tmp/prof
cap() {  # name regex cmd...
  name=$1; rx=$2; shift 2
  "$@" > gpurun_out/r2s_plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o /tmp/prof/$name -f "$@" > gpurun_out/r2s_ncu_$name.log 2>&1
  ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > gpurun_out/r02_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page details > gpurun_out/r02_${name}_details.txt 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page source --csv > gpurun_out/r02_${name}_source.csv 2>/dev/null
}
cap tile_rot fullgrid_tile_kernel python scripts/prof_calls.py C2 4096 3
cap window fullgrid_window_kernel python scripts/prof_calls.py C2 4096 3
cap roq_wide roq_kernel python scripts/prof_calls.py C3 65536 3 roq
cap relbin relbin_kernel python scripts/prof_calls.py C4 1048576 3 relbin
python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c2.csv python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1
OUT=gpurun_out/r02_stress_tail_1e7_final.txt; rm -f $OUT
python scripts/stress_tail.py --walkers 10000000 --modes mufu --out $OUT --tag "final build, 1e7 per approximant" | grep -E "^C2_small"
python scripts/stress_tail.py --walkers 2000000 --modes mufu --config C1 --approx TaylorF2_3.5PN --out $OUT --tag "C1" | grep -E "^C1"
python scripts/stress_tail.py --walkers 1000000 --modes mufu --config C2 --approx TaylorF2_3.5PN --out $OUT --tag "C2 full size" | grep -E "^C2 "
du -sh gpurun_out
