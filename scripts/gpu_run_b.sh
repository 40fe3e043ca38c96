python scripts/diag_tail.py --approx TaylorF2_3.5PN --mode mufu > gpurun_out/r2b_diag_mufu.txt 2>&1
BAJES_B200_LIB=bajes_b200/csrc/libbajes_b200_exact.so python scripts/diag_tail.py --approx TaylorF2_3.5PN --mode mufu > gpurun_out/r2b_diag_exact.txt 2>&1
BAJES_B200_LIB=bajes_b200/csrc/libbajes_b200_exact.so python scripts/stress_tail.py --walkers 1000000 --modes mufu --approx TaylorF2_3.5PN,TaylorF2_5.5PN --out gpurun_out/r2b_stress_exact.txt --tag exact-phasor
BAJES_B200_NO_ROT=1 BAJES_B200_LIB=bajes_b200/csrc/libbajes_b200_exact.so python scripts/stress_tail.py --walkers 1000000 --modes mufu --approx TaylorF2_3.5PN,TaylorF2_5.5PN --out gpurun_out/r2b_stress_exact_norot.txt --tag exact-phasor-norot
tail -5 gpurun_out/r2b_diag_mufu.txt
