#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_patch -c 4 -f -o gpurun_out/prof_convres \
  python scripts/prof_conv_res.py > gpurun_out/prof_convres.log 2>&1
echo "ncu exit $?"; grep -v PROF gpurun_out/prof_convres.log | tail -n 5
