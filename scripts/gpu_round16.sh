#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -n 3; grep -E "^FAILED|Error|watchdog" gpurun_out/pytest_gpu.log | head -n 20
for m in 1 2 0; do
  GP_STATS=$m timeout 900 python bench.py --no-cpu-baseline --ops-json gpurun_out/ops_stats$m.json > gpurun_out/bench_stats$m.log 2> gpurun_out/bench_stats$m.err
  echo "GP_STATS=$m: $(tail -n 1 gpurun_out/bench_stats$m.log | cut -c1-150)"
done
