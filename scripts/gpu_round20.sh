#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -k attention -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_attn.log 2>&1
echo "attention tests exit $?"; tail -n 3 gpurun_out/pytest_attn.log
python scripts/fattn_trace.py > gpurun_out/fattn_trace.log 2>&1; head -n 40 gpurun_out/fattn_trace.log
bash scripts/gpu_quick.sh
for cfg in "GP_FATTN_POLY=1" "GP_FATTN_V1=1"; do
  env $cfg timeout 600 python bench.py --no-cpu-baseline --ops-json gpurun_out/ops_ab.json > gpurun_out/bench_ab.log 2>&1
  echo "$cfg:"; tail -n 1 gpurun_out/bench_ab.log | cut -c1-160
  python - <<'PY'
import json
ops = json.load(open("gpurun_out/ops_ab.json"))
print("   fattn ms", sum(o["usec"] for o in ops if "fattn" in o["name"]) / 1000, "total", sum(o["usec"] for o in ops) / 1000)
PY
done
