#!/bin/bash
# Round 2, call I: residual L2 prefetch A/B; sanitizer restricted to this repo's kernels (new kernels of the round).
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multistep.py -q --no-header -p no:cacheprovider -x > $O/r2i_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2i_tests.log | cut -c1-200
for cfg in "pf" "nopf GP_NO_RES_PREFETCH=1" "pf2" "nopf2 GP_NO_RES_PREFETCH=1"; do
  set -- $cfg; tag=$1; shift
  env "$@" timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2i_ops_$tag.json > $O/r2i_bench_$tag.log 2> $O/r2i_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2i_bench_$tag.log | cut -c1-200
done
K="tests/test_gpu_kernels.py"
SEL=("$K::test_igemm_conv3x3[shape1]" "$K::test_igemm_conv3x3_residual" "$K::test_igemm_conv3x3_patch_mode_residual_relu"
     "$K::test_groupnorm_fused_into_conv3x3[1-case0]" "$K::test_groupnorm_fused_into_conv3x3[1-case4]" "$K::test_groupnorm_fused_into_conv3x3[1-case7]"
     "$K::test_attention[2-256-5-64]" "$K::test_groupnorm[shape0-32-True]" "$K::test_layernorm[100-320]"
     "tests/test_gpu_e2e.py::test_high_precision_mode_meets_the_stated_tolerance" "tests/test_gpu_multistep.py::test_ensemble_depth_matches_the_reference_function")
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "${SEL[@]}" -q --no-header -p no:cacheprovider > $O/r2i_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" $O/r2i_memcheck.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=gp --print-limit 20 python -m pytest "${SEL[@]:0:9}" -q --no-header -p no:cacheprovider > $O/r2i_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/r2i_racecheck.log | tail -3; grep -E "Write Thread|Read Thread" $O/r2i_racecheck.log | sed -E 's/.* at ([A-Za-z_0-9:<>, ()]+)\+.*/\1/' | cut -c1-90 | sort | uniq -c | head
