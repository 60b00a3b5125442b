#!/bin/bash
bash scripts/gpu_quick.sh
cp gpurun_out/ops.json gpurun_out/ops_staged.json; cp gpurun_out/bench.log gpurun_out/bench_staged.log
echo "=== A/B: direct epilogue (no staged stores, no fused statistics)"
GP_DIRECT_EPILOGUE=1 timeout 900 python bench.py --no-cpu-baseline --ops-json gpurun_out/ops_direct.json > gpurun_out/bench_direct.log 2> gpurun_out/bench_direct.err
tail -n 1 gpurun_out/bench_direct.log | cut -c1-200
timeout 300 python scripts/bench_convs.py > gpurun_out/bench_convs.log 2>&1; tail -n 14 gpurun_out/bench_convs.log
