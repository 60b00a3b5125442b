#!/bin/bash
# BASELINE.json configs beyond the headline one: DPT readout, bf16, batch-1 resolution sweep.
mkdir -p gpurun_out
: > gpurun_out/matrix.jsonl
run() { echo "== $*"; timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2> gpurun_out/matrix.err | tail -n 1 >> gpurun_out/matrix.jsonl; tail -n 1 gpurun_out/matrix.jsonl | cut -c1-260; }
run --readout dpt
run --dtype bf16
run --res 384 --batch 1
run --res 512 --batch 1
run --res 768 --batch 1
run --res 1024 --batch 1
run --cuda-graph
run --res 768 --batch 1 --cuda-graph
run --res 384 --batch 1 --cuda-graph
