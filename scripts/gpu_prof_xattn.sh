#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:xattn2 -c 3 -f -o gpurun_out/prof_xattn \
  python scripts/prof_step.py > gpurun_out/prof_xattn.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/prof_xattn.log
