#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -n 3; grep -E "^FAILED|Error|watchdog" gpurun_out/pytest_gpu.log | head -n 20
timeout 900 python bench.py --no-cpu-baseline --ops-json gpurun_out/ops.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-400; tail -n 5 gpurun_out/bench.err
timeout 900 python bench.py --no-cpu-baseline --cuda-graph > gpurun_out/bench_graph.log 2> gpurun_out/bench_graph.err
echo "bench graph exit $?"; tail -n 1 gpurun_out/bench_graph.log | cut -c1-400; tail -n 5 gpurun_out/bench_graph.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fattn_kernel -c 2 -f -o gpurun_out/prof_fattn \
  python scripts/prof_attn.py > gpurun_out/prof_attn.log 2>&1
echo "ncu fattn exit $?"; tail -n 4 gpurun_out/prof_attn.log
