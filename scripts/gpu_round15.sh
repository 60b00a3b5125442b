#!/bin/bash
bash scripts/gpu_quick.sh
timeout 300 python scripts/bench_convs.py > gpurun_out/bench_convs.log 2>&1; head -n 14 gpurun_out/bench_convs.log
