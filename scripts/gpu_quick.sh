#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -n 3; grep -E "^FAILED|Error|watchdog" gpurun_out/pytest_gpu.log | head -n 20
timeout 900 python bench.py --no-cpu-baseline --ops-json gpurun_out/ops.json $BENCH_ARGS > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-300; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json, collections
ops = json.load(open('gpurun_out/ops.json'))
tot = sum(o['usec'] for o in ops)
def cat(n):
    if 'fattn' in n: return 'fattn'
    if '.softmax' in n: return 'softmax'
    if n.endswith('.qk') or n.endswith('.pv') or 'to_vT' in n or 'to_qk' in n: return 'attn.gemm'
    if ('.norm1' in n and 'transformer' in n) or '.norm3' in n: return 'layernorm'
    if 'norm' in n: return 'groupnorm'
    if 'geglu' in n: return 'geglu'
    if 'attn2' in n: return 'xattn'
    if 'ff.' in n or 'proj_' in n or 'to_out' in n: return 'linear'
    return 'conv'
d = collections.defaultdict(lambda: [0,0,0,0])
for o in ops:
    c = cat(o['name']); d[c][0]+=o['usec']; d[c][1]+=o['flops']; d[c][2]+=o['bytes']; d[c][3]+=1
print("total ms", tot/1000)
for k,v in sorted(d.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:12s} {v[0]/1000:8.2f} ms {100*v[0]/tot:5.1f}%  n={v[3]:4d}  {v[1]/max(v[0],1e-9)/1e6:8.1f} TF/s  {v[2]/max(v[0],1e-9)/1e3:8.1f} GB/s")
PY
