#!/bin/bash
# Round 2, call L: eight epilogue warps: kernel tests, epilogue probe, bench.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q --no-header -p no:cacheprovider -x > $O/r2l_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2l_tests.log | cut -c1-200
timeout 900 python scripts/epilogue_probe.py > $O/r2l_probe.txt 2>&1; cat $O/r2l_probe.txt
for tag in a b; do
  timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2l_ops_$tag.json > $O/r2l_bench_$tag.log 2> $O/r2l_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2l_bench_$tag.log | cut -c1-200
done
