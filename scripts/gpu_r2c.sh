#!/bin/bash
# Round 2, call C: GroupNorm-fused patch kernel bring-up (per-kernel tests first, under a short timeout), then the whole
# GPU suite and the bench line.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --no-header -p no:cacheprovider -s -k "groupnorm_fused or patch_mode or igemm_conv3x3" > $O/r2c_kernels.log 2>&1
echo "kernels exit $?"; grep -E "gn\+conv|passed|failed|Error|watchdog" $O/r2c_kernels.log | cut -c1-220 | head -40
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s -x > $O/r2c_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" $O/r2c_pytest.log | tail -3; grep -E "^  [a-z]|precision=|high:|^FAILED|^E  " $O/r2c_pytest.log | cut -c1-200 | head -80
timeout 900 python bench.py --ops-json $O/r2c_ops.json > $O/r2c_bench.log 2> $O/r2c_bench.err
echo "bench exit $?"; tail -n 1 $O/r2c_bench.log | cut -c1-600
GP_NO_GN_FUSE=1 timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2c_ops_nofuse.json > $O/r2c_bench_nofuse.log 2> $O/r2c_bench_nofuse.err
echo "bench (no fuse) exit $?"; tail -n 1 $O/r2c_bench_nofuse.log | cut -c1-300
