#!/bin/bash
# Full GPU pass: parity tests, bench line, ncu launch list of the bench command, ncu --set full of the top kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --ops-json gpurun_out/ops.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -n 3 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -c 12 -f -o gpurun_out/prof_igemm \
  python scripts/prof_conv.py > gpurun_out/prof_conv.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
