#!/bin/bash
# First-contact run on the B200 box: each pytest file in its own process (a CUDA fault is sticky),
# verbose output into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for t in "tests/test_gpu_kernels.py -k direct_conv" "tests/test_gpu_kernels.py -k igemm_linear" \
         "tests/test_gpu_kernels.py -k igemm_conv3x3" "tests/test_gpu_kernels.py -k 'groupnorm or layernorm or bilinear'" \
         "tests/test_gpu_kernels.py -k attention" "tests/test_gpu_e2e.py"; do
  name=$(echo "$t" | tr ' /' '__' | tr -d "'")
  echo "=== $t"
  eval timeout 900 python -m pytest $t -m gpu -q -s -x --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?"; tail -n 25 gpurun_out/$name.log
done
timeout 300 python scripts/bench_convs.py > gpurun_out/bench_convs.log 2>&1; tail -n 30 gpurun_out/bench_convs.log
