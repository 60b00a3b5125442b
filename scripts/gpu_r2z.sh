#!/bin/bash
# Round 2, call Z: compute-sanitizer at HEAD over tests that exercise the kernels changed late in the round (eight / four
# epilogue warps, compile-time epilogue variants, pipelined row softmax, shared-memory xattn2, tile-shape model).
mkdir -p gpurun_out
O=gpurun_out
K="tests/test_gpu_kernels.py"
SEL=("$K::test_igemm_linear[1000-320-640]" "$K::test_igemm_linear_bias_residual_relu" "$K::test_igemm_conv3x3[shape1]"
     "$K::test_igemm_conv3x3_patch_mode_residual_relu" "$K::test_attention[1-2304-1-512]" "tests/test_gpu_e2e.py::test_vae_readout_matches_golden_and_oracle")
timeout 75 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "${SEL[@]}" -q --no-header -p no:cacheprovider > $O/r2z_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" $O/r2z_memcheck.log | tail -3
timeout 60 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=gp --print-limit 20 python -m pytest "${SEL[@]:0:5}" -q --no-header -p no:cacheprovider > $O/r2z_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/r2z_racecheck.log | tail -3
