#!/bin/bash
# Round 2, call P: tile model at 7 TB/s (batch 8 + batch-1 sweep), staged GEGLU and GP_STATS=2 re-checked with 8 epilogue warps.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q --no-header -p no:cacheprovider -x > $O/r2p_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2p_tests.log | cut -c1-200
run() { tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --ops-json $O/r2p_ops_$tag.json > $O/r2p_bench_$tag.log 2> $O/r2p_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2p_bench_$tag.log | cut -c1-150
}
run warm GP_NOP=1
run base GP_NOP=1
run geglu GP_STAGED_GEGLU=1
run stats2 GP_STATS=2
run nomodel GP_TILE_MODEL=0
run base2 GP_NOP=1
for tag in nomodel model; do
  if [ $tag = nomodel ]; then export GP_TILE_MODEL=0; else unset GP_TILE_MODEL; fi
  timeout 600 python bench.py --config 5 --no-cpu-baseline > $O/r2p_b1_$tag.log 2> $O/r2p_b1_$tag.err
  echo "b1 $tag exit $?"; python - <<PY
import json
d=json.loads(open("$O/r2p_b1_$tag.log").read().strip().splitlines()[-1])
sw=d.get("sweep") or d["config"].get("sweep")
print([(x["res"], round(x["ms_per_image"],2)) for x in sw])
PY
done
