#!/bin/bash
# Round 2, call D: faster GN transform (per-row patch loads, f16x2 tanh), boundary tests, full-size parity, bench A/B.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --no-header -p no:cacheprovider -x -k "groupnorm_fused or patch_mode or igemm_conv3x3" > $O/r2d_kernels.log 2>&1
echo "kernels exit $?"; grep -E "passed|failed|Error|watchdog" $O/r2d_kernels.log | cut -c1-220 | head
timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2d_ops.json > $O/r2d_bench.log 2> $O/r2d_bench.err
echo "bench exit $?"; tail -n 1 $O/r2d_bench.log | cut -c1-200
GP_PATCH_TANH32=1 timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2d_ops_tanh32.json > $O/r2d_bench_tanh32.log 2> $O/r2d_bench_tanh32.err
echo "bench (tanh32) exit $?"; tail -n 1 $O/r2d_bench_tanh32.log | cut -c1-200
GP_NO_GN_FUSE=1 timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2d_ops_nofuse.json > $O/r2d_bench_nofuse.log 2> $O/r2d_bench_nofuse.err
echo "bench (no fuse) exit $?"; tail -n 1 $O/r2d_bench_nofuse.log | cut -c1-200
timeout 1200 python -m pytest tests/test_gpu_boundary.py -q --no-header -p no:cacheprovider -s > $O/r2d_boundary.log 2>&1
echo "boundary exit $?"; grep -E "max\|err\||passed|failed|^E  |^FAILED|-> map" $O/r2d_boundary.log | cut -c1-200 | head -50
timeout 2400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_e2e.py -q --no-header -p no:cacheprovider -s > $O/r2d_parity.log 2>&1
echo "parity exit $?"; grep -E "^  [a-z]|precision=|high:|passed|failed|^FAILED|^E  " $O/r2d_parity.log | cut -c1-200 | head -80
