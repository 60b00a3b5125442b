#!/bin/bash
# Round 2, call U (gpurun --gpus 2): the 2-GPU NCCL test and the weak-scaling lines (engine arm + reference arm), as the driver launches them.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/r2u_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parallel.py -q --no-header -p no:cacheprovider > $O/r2u_tests.log 2>&1
echo "parallel test exit $?"; tail -3 $O/r2u_tests.log | cut -c1-200
: > $O/scale.jsonl
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r2u_n1.err | tail -n 1 | tee -a $O/scale.jsonl | cut -c1-160
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r2u_n2.err | grep '^{' | tail -n 1 | tee -a $O/scale.jsonl | cut -c1-160
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus 2 --config 3 --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r2u_n2g.err | grep '^{' | tail -n 1 | tee -a $O/scale.jsonl | cut -c1-160
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2> $O/r2u_ref.err | grep '^{' | tail -n 1 | tee $O/r2u_ref.log | cut -c1-200
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/scale.jsonl") if l.startswith("{")]
base = rows[0]["value"]
for r in rows:
    print(f"N={r['n_gpus']} {r['config']['workload'][:40]}: {r['value']:.1f} images/s ({r['value'] / base / r['n_gpus'] * 100:.1f} % of linear) e2e {r['e2e']['value']:.1f}")
PY
