#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -n 3; grep -E "^FAILED|^E  |token context|empty prompt" gpurun_out/pytest_gpu.log | head -n 30
