#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -c 17 -f -o gpurun_out/prof_igemm \
  python scripts/prof_conv.py > gpurun_out/prof_conv.log 2>&1
echo "ncu igemm exit $?"; tail -n 4 gpurun_out/prof_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fattn_kernel -c 2 -f -o gpurun_out/prof_fattn \
  python scripts/prof_attn.py > gpurun_out/prof_attn.log 2>&1
echo "ncu fattn exit $?"; tail -n 3 gpurun_out/prof_attn.log
