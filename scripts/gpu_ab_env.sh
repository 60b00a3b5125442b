#!/bin/bash
# usage: gpu_ab_env.sh "ENV1=a" "ENV2=b ENV3=c" ...   -> bench (6 steps) per setting, fattn / total ms from the per-op pass
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --ops-json gpurun_out/ops_ab.json > gpurun_out/bench_ab.log 2>&1
  echo "$cfg: $(tail -n 1 gpurun_out/bench_ab.log | cut -c1-75)"
  python - <<'PY'
import json
ops = json.load(open("gpurun_out/ops_ab.json"))
print("   fattn ms", round(sum(o["usec"] for o in ops if "fattn" in o["name"]) / 1000, 2), "total", round(sum(o["usec"] for o in ops) / 1000, 2))
PY
done
