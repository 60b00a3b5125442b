#!/bin/bash
# Round 2, call V: four epilogue warps where eight cost a pipeline stage (long-K BN = 256 layers): full suite, A/B, refreshed bench line.
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > $O/r2v_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2v_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2v_smoke.log 2>&1; tail -2 $O/r2v_smoke.log
run() { tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --ops-json $O/r2v_ops_$tag.json > $O/r2v_bench_$tag.log 2> $O/r2v_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2v_bench_$tag.log | cut -c1-150
}
run warm GP_NOP=1
run base GP_NOP=1
run epi8 GP_EPI_WARPS8=1
run base2 GP_NOP=1
run epi8b GP_EPI_WARPS8=1
timeout 900 python bench.py --ops-json $O/ops.json > $O/bench.log 2> $O/bench.err
echo "full bench exit $?"; tail -n 1 $O/bench.log | cut -c1-200
