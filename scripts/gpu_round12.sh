#!/bin/bash
bash scripts/gpu_quick.sh
echo "=== res 1024 batch 1"
timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3 --res 1024 --batch 1 > gpurun_out/b1024.log 2> gpurun_out/b1024.err
echo "exit $?"; tail -n 1 gpurun_out/b1024.log | cut -c1-200; tail -n 8 gpurun_out/b1024.err
