# final evidence of round 2 (one GPU): launch lists, ncu captures of the node-axis kernels, ROQ basis bench, bench line
set -x
mkdir -p /tmp/prof
python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_compressed.csv python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1
cap() {  # name regex cmd...
  name=$1; rx=$2; shift 2
  "$@" > gpurun_out/r2f_plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o /tmp/prof/$name -f "$@" > gpurun_out/r2f_ncu_$name.log 2>&1
  ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > gpurun_out/r02_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page details > gpurun_out/r02_${name}_details.txt 2>/dev/null
  ncu -i /tmp/prof/$name.ncu-rep --page source --csv > gpurun_out/r02_${name}_source.csv 2>/dev/null
}
cap tile_nodes fullgrid_tile_kernel python scripts/prof_calls.py C2 4096 3
cap window_nodes fullgrid_window_kernel python scripts/prof_calls.py C2 4096 3
cap classify classify_kernel python scripts/prof_calls.py C2 4096 3
python scripts/roq_basis_bench.py 8192 > gpurun_out/r02_roq_basis_bench.json 2> gpurun_out/r02_roq_basis_bench.err; tail -2 gpurun_out/r02_roq_basis_bench.err
cap rb_project rb_project_kernel python scripts/roq_basis_bench.py 4096
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -3 gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
du -sh gpurun_out
