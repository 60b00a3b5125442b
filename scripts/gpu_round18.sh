#!/bin/bash
bash scripts/gpu_quick.sh
python scripts/fattn_trace.py > gpurun_out/fattn_trace.log 2>&1; head -n 12 gpurun_out/fattn_trace.log; grep -A8 "MMA issuer" gpurun_out/fattn_trace.log | head -n 10
timeout 300 python scripts/bench_convs.py > gpurun_out/bench_convs.log 2>&1; head -n 14 gpurun_out/bench_convs.log
