#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -k attention -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_attn.log 2>&1
echo "attention tests exit $?"; tail -n 2 gpurun_out/pytest_attn.log
for cfg in "GP_X=0" "GP_FATTN_PP=1" "GP_FATTN_STAGGER=1500" "GP_FATTN_STAGGER=2500" "GP_FATTN_PP=1 GP_FATTN_POLY=1" "GP_FATTN_STAGGER=1500 GP_FATTN_POLY=1" "GP_FATTN_POLY=1"; do
  env $cfg timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --ops-json gpurun_out/ops_ab.json > gpurun_out/bench_ab.log 2>&1
  echo "$cfg: $(tail -n 1 gpurun_out/bench_ab.log | cut -c1-75)"
  python - <<'PY'
import json
ops = json.load(open("gpurun_out/ops_ab.json"))
print("   fattn ms", round(sum(o["usec"] for o in ops if "fattn" in o["name"]) / 1000, 2), "total", round(sum(o["usec"] for o in ops) / 1000, 2))
PY
done
GP_FATTN_PP=1 python scripts/fattn_trace.py > gpurun_out/fattn_trace_pp.log 2>&1; head -n 14 gpurun_out/fattn_trace_pp.log
GP_FATTN_STAGGER=1500 python scripts/fattn_trace.py > gpurun_out/fattn_trace_st.log 2>&1; head -n 14 gpurun_out/fattn_trace_st.log
