mkdir -p /tmp/prof
cap() {  # name regex cmd...
  name=$1; rx=$2; shift 2
  "$@" > gpurun_out/r2n_plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o /tmp/prof/$name -f "$@" > gpurun_out/r2n_ncu_$name.log 2>&1
  ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > gpurun_out/r02_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page details > gpurun_out/r02_${name}_details.txt 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page source --csv > gpurun_out/r02_${name}_source.csv 2>/dev/null
  ls -la /tmp/prof/$name.ncu-rep
}
cap tile_rot fullgrid_tile_kernel python scripts/prof_calls.py C2 4096 3
cap window fullgrid_window_kernel python scripts/prof_calls.py C2 4096 3
cap roq_wide roq_kernel python scripts/prof_calls.py C3 65536 3 roq
cap relbin relbin_kernel python scripts/prof_calls.py C4 1048576 3 relbin
du -sh gpurun_out
