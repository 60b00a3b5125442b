#!/bin/bash
mkdir -p gpurun_out
for st in 4 3 2; do
  echo "GP_PATCH_STAGES=$st"
  GP_PATCH_STAGES=$st timeout 300 python scripts/bench_convs.py 2>&1 | head -n 2
done
