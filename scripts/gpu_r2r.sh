#!/bin/bash
# Round 2, call R: compile-time specialised staged epilogue + whole-bias-in-smem: tests, short-K linears alone, bench.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q --no-header -p no:cacheprovider -x > $O/r2r_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2r_tests.log | cut -c1-200
GP_PROF_ITERS=20 timeout 300 python scripts/prof_linear.py > $O/r2r_lin_alone.txt 2>&1; cat $O/r2r_lin_alone.txt
GP_PROF_ITERS=20 GP_NO_BIAS_ALL=1 timeout 300 python scripts/prof_linear.py > $O/r2r_lin_nobiasall.txt 2>&1; cat $O/r2r_lin_nobiasall.txt
GP_PROF_ITERS=20 GP_BENCH_RES=1 timeout 300 python scripts/prof_linear.py > $O/r2r_lin_res.txt 2>&1; cat $O/r2r_lin_res.txt
timeout 300 python scripts/epilogue_probe.py > $O/r2r_probe.txt 2>&1; cat $O/r2r_probe.txt
run() { tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --ops-json $O/r2r_ops_$tag.json > $O/r2r_bench_$tag.log 2> $O/r2r_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2r_bench_$tag.log | cut -c1-150
}
run warm GP_NOP=1
run base GP_NOP=1
run nobiasall GP_NO_BIAS_ALL=1
run base2 GP_NOP=1
