#!/bin/bash
# Round-end measurement pass: tests, the contract bench line (with cpu_baseline), the reference arm, the
# ncu launch list of the same bench command, clocks.  Summarise with scripts/make_profiles.py <tag>.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit,temperature.gpu --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -n 2
timeout 900 python bench.py --ops-json gpurun_out/ops.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
echo "ref exit $?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
