#!/bin/bash
# DRAM bytes of every launch of one warm step (-> profiles/<tag>_step_dram.json via make_profiles.py)
mkdir -p gpurun_out
timeout 1200 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none --csv --log-file gpurun_out/step_dram.csv python scripts/prof_step.py > gpurun_out/prof_step.log 2>&1
echo "ncu step exit $?"; wc -l gpurun_out/step_dram.csv; tail -n 2 gpurun_out/prof_step.log
