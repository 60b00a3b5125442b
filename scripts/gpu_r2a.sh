#!/bin/bash
# Round 2, call A: baseline parity at the BASELINE shapes (default precision), diffusers probe,
# compute-sanitizer memcheck / racecheck over a set of per-kernel tests.
mkdir -p gpurun_out
O=gpurun_out
{ python -c "import diffusers; print('diffusers', diffusers.__version__)"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "diffusers|accelerate|xformers|peft|omegaconf"; echo "wheelhouse grep exit $?"; } > $O/r2a_diffusers_probe.txt 2>&1
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -q --no-header -p no:cacheprovider -s > $O/r2a_fullsize.log 2>&1
echo "fullsize exit $?"; grep -E "^  [a-z]|precision=|passed|failed|Error" $O/r2a_fullsize.log | head -60
K="tests/test_gpu_kernels.py"
SEL=("$K::test_igemm_linear[256-64-64]" "$K::test_igemm_conv3x3[shape0]" "$K::test_igemm_conv3x3[shape1]" "$K::test_igemm_conv3x3_residual"
     "$K::test_igemm_conv3x3_upsample_fused[shape0]" "$K::test_igemm_conv3x3_stride2[shape0-1]" "$K::test_igemm_conv3x3_patch_mode_residual_relu"
     "$K::test_attention[2-256-5-64]" "$K::test_attention[1-1024-1-512]" "$K::test_groupnorm[shape0-32-True]" "$K::test_layernorm[100-320]"
     "tests/test_gpu_e2e.py::test_vae_readout_matches_golden_and_oracle")
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "${SEL[@]}" -q --no-header -p no:cacheprovider > $O/r2a_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" $O/r2a_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -m pytest "${SEL[@]:0:11}" -q --no-header -p no:cacheprovider > $O/r2a_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/r2a_racecheck.log | tail -3
