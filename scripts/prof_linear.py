"""Short-K linears of the UNet's first level, once each, for `ncu --set full -k regex:igemm_kernel` (source-level stalls),
then timed alone under a few A/B switches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpercept_b200 import engine as E  # noqa: E402

SHAPES = [(8, 96, 96, 320, 2560, 1, 0), (8, 96, 96, 320, 320, 1, 0), (8, 96, 96, 1280, 320, 1, 0), (8, 96, 96, 320, 640, 1, 0)]
iters = int(os.environ.get("GP_PROF_ITERS", "1"))
for s in SHAPES:
    us, fl = E.bench_conv(torch.float16, *s, iters=iters)
    print(s, f"{us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s", flush=True)
