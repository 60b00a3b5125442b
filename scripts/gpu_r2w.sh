#!/bin/bash
# Round 2, call W: patch-resident layers without a residual on four epilogue warps + the fourth weight-ring stage: tests + A/B.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q --no-header -p no:cacheprovider -x > $O/r2w_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2w_tests.log | cut -c1-200
run() { tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --ops-json $O/r2w_ops_$tag.json > $O/r2w_bench_$tag.log 2> $O/r2w_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2w_bench_$tag.log | cut -c1-150
}
run warm GP_NOP=1
run base GP_NOP=1
run ne8 GP_PATCH_NE8=1
run base2 GP_NOP=1
run ne8b GP_PATCH_NE8=1
