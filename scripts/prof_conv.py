"""Launches the dominant implicit-GEMM shapes once each (for `ncu --set full -k regex:igemm_kernel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpercept_b200 import engine as E  # noqa: E402

for s in [(8, 768, 768, 128, 128, 3, 0), (8, 384, 384, 256, 256, 3, 0), (8, 96, 96, 320, 2560, 1, 0)]:
    us, fl = E.bench_conv(torch.float16, *s, iters=1)
    print(s, us, fl / us / 1e6)
