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
# the .ncu-rep files (17 launches with --set full + source) exceed gpurun's 64 MiB return limit: summarise on the box, drop them
mkdir -p gpurun_out/profiles_out
GP_PROFILES_DIR=gpurun_out/profiles_out python scripts/make_profiles.py r2 > gpurun_out/make_profiles.log 2>&1; tail -n 2 gpurun_out/make_profiles.log | cut -c1-200
python scripts/ncu_source_summary.py gpurun_out/prof_igemm.ncu-rep 11 5760 > gpurun_out/profiles_out/r2_linear_after_ncu_source.txt 2>&1
python scripts/ncu_source_summary.py gpurun_out/prof_igemm.ncu-rep 3 0 > gpurun_out/profiles_out/r2_conv128_ncu_source.txt 2>&1
python scripts/ncu_source_summary.py gpurun_out/prof_fattn.ncu-rep 1 0 > gpurun_out/profiles_out/r2_fattn_ncu_source.txt 2>&1
rm -f gpurun_out/prof_igemm.ncu-rep gpurun_out/prof_fattn.ncu-rep
du -sm gpurun_out
