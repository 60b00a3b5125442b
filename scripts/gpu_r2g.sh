#!/bin/bash
# Round 2, call G: transform-loop variants of the patch kernel (trace + bench), f4 / boundary tests.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multistep.py tests/test_gpu_boundary.py -q --no-header -p no:cacheprovider -s > $O/r2g_f4.log 2>&1
echo "f4/boundary exit $?"; grep -E "max\|err\||passed|failed|^E  |^FAILED" $O/r2g_f4.log | cut -c1-200 | head -40
{ for m in 0 1 2 3; do for t in 0 1; do
    echo "#### GP_PATCH_XFORM=$m tanh32=$t"
    if [ $t = 1 ]; then export GP_PATCH_TANH32=1; else unset GP_PATCH_TANH32; fi
    GP_PATCH_XFORM=$m python scripts/patch_trace.py 128 128 768 768 8 | sed -n '1,2p;8,14p'
  done; done; unset GP_PATCH_TANH32; } > $O/r2g_trace.log 2>&1
cat $O/r2g_trace.log
for cfg in "nofuse GP_NO_GN_FUSE=1" "x0 GP_PATCH_XFORM=0" "x2t GP_PATCH_XFORM=2 GP_PATCH_TANH32=1" "x1 GP_PATCH_XFORM=1"; do
  set -- $cfg; tag=$1; shift
  env "$@" timeout 900 python bench.py --no-cpu-baseline --ops-json $O/r2g_ops_$tag.json > $O/r2g_bench_$tag.log 2> $O/r2g_bench_$tag.err
  echo "bench $tag exit $?"; tail -n 1 $O/r2g_bench_$tag.log | cut -c1-200
done
