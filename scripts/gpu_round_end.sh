#!/bin/bash
# Everything the round-end profiles need, in one box session.
bash scripts/gpu_final.sh
bash scripts/gpu_matrix.sh
bash scripts/gpu_round23.sh
bash scripts/gpu_prof.sh
