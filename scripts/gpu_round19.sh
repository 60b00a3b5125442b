#!/bin/bash
bash scripts/gpu_quick.sh
GP_FATTN_POLY=0 timeout 600 python bench.py --no-cpu-baseline --ops-json gpurun_out/ops_nopoly.json > gpurun_out/bench_nopoly.log 2>&1
echo "nopoly:"; tail -n 1 gpurun_out/bench_nopoly.log | cut -c1-200
python - <<'PY'
import json
for f in ("ops.json", "ops_nopoly.json"):
    ops = json.load(open("gpurun_out/" + f))
    print(f, "fattn ms", sum(o["usec"] for o in ops if "fattn" in o["name"]) / 1000, "total", sum(o["usec"] for o in ops) / 1000)
PY
python scripts/fattn_trace.py > gpurun_out/fattn_trace.log 2>&1; head -n 12 gpurun_out/fattn_trace.log; grep -A8 "MMA issuer" gpurun_out/fattn_trace.log | head -n 10
timeout 300 python scripts/bench_convs.py > gpurun_out/bench_convs.log 2>&1; head -n 14 gpurun_out/bench_convs.log
timeout 120 build/exp_ex2 > gpurun_out/exp_ex2.log 2>&1; cat gpurun_out/exp_ex2.log
