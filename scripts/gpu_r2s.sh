#!/bin/bash
# Round 2, call S: full GPU suite + smoke at HEAD, bench A/B of nothing (two runs), GEGLU per-op check.
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > $O/r2s_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2s_tests.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2s_smoke.log 2>&1; tail -2 $O/r2s_smoke.log
for tag in a b; do
  timeout 600 python bench.py --no-cpu-baseline --ops-json $O/r2s_ops_$tag.json > $O/r2s_bench_$tag.log 2> $O/r2s_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2s_bench_$tag.log | cut -c1-150
done
GP_PROF_ITERS=20 timeout 300 python scripts/prof_linear.py > $O/r2s_lin_alone.txt 2>&1; cat $O/r2s_lin_alone.txt
