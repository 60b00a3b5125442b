#!/bin/bash
# 2-GPU weak-scaling check: the N=1 line and the N=2 torchrun line back to back on the same box.
mkdir -p gpurun_out
rm -f gpurun_out/scale.jsonl
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.log 2> gpurun_out/scale_n1.err
echo "n1 exit $?"; tail -n 1 gpurun_out/scale_n1.log | tee -a gpurun_out/scale.jsonl | cut -c1-260
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n2.log 2> gpurun_out/scale_n2.err
echo "n2 exit $?"; tail -n 1 gpurun_out/scale_n2.log | tee -a gpurun_out/scale.jsonl | cut -c1-260; tail -n 5 gpurun_out/scale_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --gather > gpurun_out/scale_n2g.log 2> gpurun_out/scale_n2g.err
echo "n2 gather exit $?"; tail -n 1 gpurun_out/scale_n2g.log | tee -a gpurun_out/scale.jsonl | cut -c1-260; tail -n 5 gpurun_out/scale_n2g.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/scale_ref2.log 2> gpurun_out/scale_ref2.err
echo "ref n2 exit $?"; tail -n 1 gpurun_out/scale_ref2.log | cut -c1-260; tail -n 3 gpurun_out/scale_ref2.err
