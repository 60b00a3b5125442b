#!/bin/bash
# Everything the round-end profiles need, in one box session.
bash scripts/gpu_final.sh
bash scripts/gpu_matrix.sh
bash scripts/gpu_step_dram.sh
bash scripts/gpu_prof.sh
