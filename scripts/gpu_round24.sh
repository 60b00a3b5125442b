#!/bin/bash
bash scripts/gpu_quick.sh
python - <<'PY'
import json
ops=json.load(open('gpurun_out/ops.json'))
for o in ops:
    n=o['name']
    if ('vae.encoder.down_blocks.0.resnets' in n or 'vae.decoder.up_blocks.3.resnets' in n or 'vae.decoder.up_blocks.2.resnets.1' in n or 'unet.down_blocks.0.resnets.0' in n) and 'conv' in n:
        print(f"{o['usec']:8.1f} us {o['flops']/max(o['usec'],1e-9)/1e6:7.0f} TF/s  {n}")
PY
