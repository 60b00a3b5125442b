#!/bin/bash
# Round 2, call F: special-function throughput microbenchmark; N-tile policy A/B (staged epilogue for Cout = 320 / 640).
mkdir -p gpurun_out
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/exp_mufu scripts/exp_mufu.cu && /tmp/exp_mufu > $O/r2f_mufu.txt 2>&1
echo "mufu exit $?"; cat $O/r2f_mufu.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py::test_vae_readout_matches_golden_and_oracle tests/test_gpu_e2e.py::test_dpt_readout_matches_golden -q --no-header -p no:cacheprovider -x > $O/r2f_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2f_tests.log
for pol in 0 1 2; do
  GP_NO_GN_FUSE=1 GP_BN_POLICY=$pol timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2f_ops_bn$pol.json > $O/r2f_bench_bn$pol.log 2> $O/r2f_bench_bn$pol.err
  echo "bench (nofuse, bn policy $pol) exit $?"; tail -n 1 $O/r2f_bench_bn$pol.log | cut -c1-330
done
timeout 1200 python -m pytest tests/test_gpu_multistep.py tests/test_gpu_boundary.py -q --no-header -p no:cacheprovider -s > $O/r2f_f4.log 2>&1
echo "f4/boundary exit $?"; grep -E "max\|err\||passed|failed|^E  |^FAILED" $O/r2f_f4.log | cut -c1-200 | head -40
