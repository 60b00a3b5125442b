#!/bin/bash
# Round 2, call B: high-precision mode bring-up (64x64 first), then call A's programme.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest "tests/test_gpu_e2e.py::test_high_precision_mode_meets_the_stated_tolerance" "tests/test_gpu_e2e.py::test_vae_readout_matches_golden_and_oracle" -q --no-header -p no:cacheprovider -s > $O/r2b_hp64.log 2>&1
echo "hp64 exit $?"; grep -E "high:|passed|failed|Error|error" $O/r2b_hp64.log | head -30
bash scripts/gpu_r2a.sh
