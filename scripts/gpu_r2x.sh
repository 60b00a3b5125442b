#!/bin/bash
# Round 2, call X: the GPU test files call W did not run + smoke, at HEAD.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_multistep.py tests/test_gpu_imgproc.py tests/test_bench_contract.py -m gpu -q --no-header -p no:cacheprovider > $O/r2x_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2x_tests.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2x_smoke.log 2>&1; tail -2 $O/r2x_smoke.log
