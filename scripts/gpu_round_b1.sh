#!/bin/bash
bash scripts/gpu_tests_only.sh
mkdir -p gpurun_out
: > gpurun_out/matrix_b1.jsonl
for r in 384 768; do
  timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --res $r --batch 1 2> gpurun_out/matrix.err | tail -n 1 >> gpurun_out/matrix_b1.jsonl
  tail -n 1 gpurun_out/matrix_b1.jsonl | cut -c1-200
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -n 1 gpurun_out/bench.log | cut -c1-200
