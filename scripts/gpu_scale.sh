#!/bin/bash
# Weak-scaling check on one box: N = 1 and every N in "$@" (default: 2 4 8), the engine arm and the reference arm,
# launched exactly as the driver launches them.  gpurun --gpus <maxN> -- 'bash scripts/gpu_scale.sh 2 4 8'
mkdir -p gpurun_out
: > gpurun_out/scale.jsonl
NS=${@:-2 4 8}
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_n1.err | tail -n 1 | tee -a gpurun_out/scale.jsonl | cut -c1-200
port=29520
for n in $NS; do
  port=$((port + 1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_n$n.err | grep '^{' | tail -n 1 | tee -a gpurun_out/scale.jsonl | cut -c1-200
done
port=$((port + 1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2> gpurun_out/scale_ref.err | grep '^{' | tail -n 1 | cut -c1-200
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/scale.jsonl") if l.startswith("{")]
base = rows[0]["value"]
for r in rows:
    print(f"N={r['n_gpus']}: {r['value']:.1f} images/s  ({r['value'] / base / r['n_gpus'] * 100:.1f} % of linear)  e2e {r['e2e']['value']:.1f}")
PY
