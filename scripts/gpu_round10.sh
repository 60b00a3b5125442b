#!/bin/bash
bash scripts/gpu_quick.sh
cp gpurun_out/ops.json gpurun_out/ops_patch.json
echo "=== A/B: GP_NO_PATCH=1"
GP_NO_PATCH=1 timeout 900 python bench.py --no-cpu-baseline --ops-json gpurun_out/ops_nopatch.json > gpurun_out/bench_nopatch.log 2> gpurun_out/bench_nopatch.err
tail -n 1 gpurun_out/bench_nopatch.log | cut -c1-200
timeout 300 python scripts/bench_convs.py > gpurun_out/bench_convs.log 2>&1; head -n 3 gpurun_out/bench_convs.log
