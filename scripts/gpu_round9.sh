#!/bin/bash
./build/exp_baseoffset > gpurun_out/exp_baseoffset.log 2>&1; cat gpurun_out/exp_baseoffset.log
bash scripts/gpu_quick.sh
