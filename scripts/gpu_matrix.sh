#!/bin/bash
# BASELINE.json configs beyond the headline one (bench.py --config): normal + gather path, DPT readout, batch-1 sweep,
# the 512x512 plumbing case, bf16 storage, the high-precision mode, forced CUDA-graph replay.
mkdir -p gpurun_out
: > gpurun_out/matrix.jsonl
run() { echo "== $*"; timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2> gpurun_out/matrix.err | tail -n 1 >> gpurun_out/matrix.jsonl; tail -n 1 gpurun_out/matrix.jsonl | cut -c1-260; }
run --config 3
run --config 4
run --config 5
run --config 1
run --dtype bf16
run --precision high --batch 2
run --cuda-graph
