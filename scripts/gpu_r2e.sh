#!/bin/bash
# Round 2, call E: per-chunk trace of the patch kernel (why the fused convs lost), arena-skew experiment, im2col stem,
# new bench contract (CPU arm at 768x768), reruns of the two fixed tests.
mkdir -p gpurun_out
O=gpurun_out
{ python scripts/patch_trace.py 128 128 768 768 8; python scripts/patch_trace.py 256 128 768 768 8; python scripts/patch_trace.py 256 256 384 384 8;
  echo "---- GP_PATCH_TANH32=1"; GP_PATCH_TANH32=1 python scripts/patch_trace.py 128 128 768 768 8; } > $O/r2e_trace.log 2>&1
echo "trace exit $?"; head -70 $O/r2e_trace.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_boundary.py::test_from_run_args_on_disk_layout_drives_the_engine "tests/test_gpu_e2e.py::test_sizes_that_are_multiples_of_8_only" tests/test_gpu_e2e.py::test_vae_readout_matches_golden_and_oracle -q --no-header -p no:cacheprovider -x > $O/r2e_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2e_tests.log
timeout 900 python bench.py --ops-json $O/r2e_ops.json > $O/r2e_bench.log 2> $O/r2e_bench.err
echo "bench exit $?"; tail -n 1 $O/r2e_bench.log | cut -c1-1800; tail -3 $O/r2e_bench.err
GP_ARENA_SKEW=132 timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2e_ops_skew.json > $O/r2e_bench_skew.log 2> $O/r2e_bench_skew.err
echo "bench (skew) exit $?"; tail -n 1 $O/r2e_bench_skew.log | cut -c1-200
GP_NO_GN_FUSE=1 timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2e_ops_nofuse.json > $O/r2e_bench_nofuse.log 2> $O/r2e_bench_nofuse.err
echo "bench (no fuse) exit $?"; tail -n 1 $O/r2e_bench_nofuse.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2e_bench_ref.log 2> $O/r2e_bench_ref.err
echo "ref exit $?"; tail -n 1 $O/r2e_bench_ref.log | cut -c1-900
