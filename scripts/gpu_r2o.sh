#!/bin/bash
# Round 2, call O: (BN, MT) tile-shape model (waves + L2 traffic), pipelined gn_apply: tests + A/B benches.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_multistep.py -q --no-header -p no:cacheprovider -x > $O/r2o_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/r2o_tests.log | cut -c1-200
run() { tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --ops-json $O/r2o_ops_$tag.json > $O/r2o_bench_$tag.log 2> $O/r2o_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2o_bench_$tag.log | cut -c1-150
}
run base GP_NOP=1
run nomodel GP_TILE_MODEL=0
run gnpipe1 GP_GN_PIPE=1
run gnpipe2 GP_GN_PIPE=2
run base2 GP_NOP=1
run nomodel2 GP_TILE_MODEL=0
for tag in model nomodel; do
  if [ $tag = nomodel ]; then export GP_TILE_MODEL=0; else unset GP_TILE_MODEL; fi
  timeout 600 python bench.py --config 5 --no-cpu-baseline > $O/r2o_b1_$tag.log 2> $O/r2o_b1_$tag.err
  echo "b1 $tag exit $?"; python - <<PY
import json
d=json.loads(open("$O/r2o_b1_$tag.log").read().strip().splitlines()[-1])
sw=d.get("sweep") or d["config"].get("sweep")
print([(x["res"], round(x["ms_per_image"],2)) for x in sw])
PY
done
