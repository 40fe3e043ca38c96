set -x
python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_c2_4096.csv python scripts/prof_calls.py C2 4096 3 > /dev/null 2>&1
python scripts/prof_calls.py C2 64 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_c2_64.csv python scripts/prof_calls.py C2 64 3 > /dev/null 2>&1
