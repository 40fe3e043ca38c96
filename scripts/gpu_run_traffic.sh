# ncu --set full captures of every dominant kernel at the FINAL kernel sources -> gpurun_out/r02_*_{raw.csv,details.txt,source.csv}
mkdir -p /tmp/prof
cap() {  # name regex cmd...
  name=$1; rx=$2; shift 2
  "$@" > gpurun_out/r2g_plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o /tmp/prof/$name -f "$@" > gpurun_out/r2g_ncu_$name.log 2>&1
  ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > gpurun_out/r02_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page details > gpurun_out/r02_${name}_details.txt 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page source --csv > gpurun_out/r02_${name}_source.csv 2>/dev/null
}
cap tile_nodes fullgrid_tile_kernel python scripts/prof_calls.py C2 4096 3
cap window_nodes fullgrid_window_kernel python scripts/prof_calls.py C2 4096 3
export BAJES_B200_COMPRESS=0
cap tile_rot fullgrid_tile_kernel python scripts/prof_calls.py C2 4096 3
cap window fullgrid_window_kernel python scripts/prof_calls.py C2 4096 3
python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fullgrid.csv python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1
unset BAJES_B200_COMPRESS
cap roq_wide roq_kernel python scripts/prof_calls.py C3 65536 3 roq
cap relbin relbin_kernel python scripts/prof_calls.py C4 1048576 3 relbin
cap rb_project rb_project_kernel python scripts/roq_basis_bench.py 4096
python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_compressed.csv python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1
