#!/bin/bash
# Round 2, call M: pipelined row softmax, shared-memory xattn2, programmatic dependent launch: tests + A/B benches.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q --no-header -p no:cacheprovider -x > $O/r2m_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2m_tests.log | cut -c1-200
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --ops-json $O/r2m_ops_$tag.json > $O/r2m_bench_$tag.log 2> $O/r2m_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2m_bench_$tag.log | cut -c1-150
}
run base GP_NOP=1
run nopdl GP_PDL=0
run old GP_PDL=0 GP_SOFTMAX_PIPE=0 GP_XATTN_SMEM=0
run ln2 GP_LN_TOK=2
run base2 GP_NOP=1
for tag in pdl nopdl; do
  if [ $tag = nopdl ]; then export GP_PDL=0; else unset GP_PDL; fi
  timeout 600 python bench.py --config 5 --no-cpu-baseline > $O/r2m_b1_$tag.log 2> $O/r2m_b1_$tag.err
  echo "b1 $tag exit $?"; tail -n 1 $O/r2m_b1_$tag.log | cut -c1-600
done
