#!/bin/bash
# Round 2, call Q: why the short-K linears sit at ~25 % tensor: ncu --set full with source, and timings alone.
mkdir -p gpurun_out
O=gpurun_out
GP_PROF_ITERS=20 timeout 300 python scripts/prof_linear.py > $O/r2q_lin_alone.txt 2>&1; cat $O/r2q_lin_alone.txt
GP_PROF_ITERS=20 GP_DIRECT_EPILOGUE=1 timeout 300 python scripts/prof_linear.py > $O/r2q_lin_direct.txt 2>&1; cat $O/r2q_lin_direct.txt
GP_PROF_ITERS=20 GP_BENCH_RES=1 timeout 300 python scripts/prof_linear.py > $O/r2q_lin_res.txt 2>&1; cat $O/r2q_lin_res.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -c 4 -f -o $O/prof_linear \
  python scripts/prof_linear.py > $O/r2q_prof.log 2>&1
echo "ncu exit $?"; tail -n 3 $O/r2q_prof.log
ls -la $O/prof_linear.ncu-rep
