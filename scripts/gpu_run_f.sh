OUT=gpurun_out/r02_stress_tail_1e7.txt
rm -f $OUT
python scripts/stress_tail.py --walkers 10000000 --modes mufu --out $OUT --tag "default build, 1e7 per approximant" | grep -E "^C2_small"
python scripts/stress_tail.py --walkers 10000000 --modes poly32 --approx TaylorF2_3.5PN,TaylorF2_5.5PN_7.5PNTides2020 --out $OUT --tag "poly32" | grep -E "^C2_small"
python scripts/stress_tail.py --walkers 2000000 --modes mufu --marg-phi --approx TaylorF2_3.5PN --out $OUT --tag "phi_ref marginalised" | grep -E "^C2_small"
python scripts/stress_tail.py --walkers 2000000 --modes mufu --config C1 --approx TaylorF2_3.5PN,TaylorF2_5.5PN --out $OUT --tag "C1" | grep -E "^C1"
python scripts/stress_tail.py --walkers 1000000 --modes mufu --config C2 --approx TaylorF2_3.5PN,TaylorF2_5.5PN_7.5PNTides2020 --out $OUT --tag "C2 full size" | grep -E "^C2 "
