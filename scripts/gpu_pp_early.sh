#!/bin/bash
mkdir -p gpurun_out
for e in 4 3 2 1 0; do
  echo "GP_FATTN_PP_EARLY=$e"
  GP_FATTN_PP_EARLY=$e python scripts/fattn_trace.py 2>&1 | sed -n 3,8p
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -k attention -q --no-header -p no:cacheprovider -x 2>&1 | tail -n 2
for e in 4 2; do
  GP_FATTN_PP_EARLY=$e timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --ops-json gpurun_out/ops_ab.json > gpurun_out/bench_ab.log 2>&1
  echo "PP_EARLY=$e: $(tail -n 1 gpurun_out/bench_ab.log | cut -c1-75)"
  python - <<'PY'
import json
ops = json.load(open("gpurun_out/ops_ab.json"))
print("   fattn ms", round(sum(o["usec"] for o in ops if "fattn" in o["name"]) / 1000, 2), "total", round(sum(o["usec"] for o in ops) / 1000, 2))
PY
done
