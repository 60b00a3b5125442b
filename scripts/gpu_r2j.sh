#!/bin/bash
# Round 2, call J: the profile set (bench + per-op JSON, reference arm, ncu launch list, config matrix, per-launch DRAM
# traffic of a step, ncu --set full captures).  Summarise with: python scripts/make_profiles.py r2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit,temperature.gpu --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python bench.py --ops-json gpurun_out/ops.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
echo "ref exit $?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches.csv
bash scripts/gpu_matrix.sh
bash scripts/gpu_step_dram.sh
bash scripts/gpu_prof.sh
