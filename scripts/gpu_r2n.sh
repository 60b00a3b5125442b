#!/bin/bash
# Round 2, call N: wave-aware N tile for small maps: full GPU suite + A/B benches (batch 8 and the batch-1 sweep).
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > $O/r2n_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2n_tests.log | cut -c1-200
run() { tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --ops-json $O/r2n_ops_$tag.json > $O/r2n_bench_$tag.log 2> $O/r2n_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2n_bench_$tag.log | cut -c1-150
}
run base GP_NOP=1
run nowaves GP_BN_WAVES=0
run base2 GP_NOP=1
for tag in waves nowaves; do
  if [ $tag = nowaves ]; then export GP_BN_WAVES=0; else unset GP_BN_WAVES; fi
  timeout 600 python bench.py --config 5 --no-cpu-baseline > $O/r2n_b1_$tag.log 2> $O/r2n_b1_$tag.err
  echo "b1 $tag exit $?"; python - <<PY
import json
d=json.loads(open("$O/r2n_b1_$tag.log").read().strip().splitlines()[-1])
sw=d.get("sweep") or d["config"].get("sweep")
print([(x["res"], round(x["ms_per_image"],2)) for x in sw])
PY
done
unset GP_BN_WAVES
timeout 600 python bench.py --config 4 --no-cpu-baseline > $O/r2n_dpt.log 2>&1; tail -n 1 $O/r2n_dpt.log | cut -c1-150
