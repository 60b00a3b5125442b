"""conv3x3 128->128 @768^2 B=8 with and without a residual input (ncu --set full -k regex:igemm_patch)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpercept_b200 import engine as E  # noqa: E402

g = torch.Generator().manual_seed(0)
x = (torch.randn((8, 768, 768, 128), generator=g) * 0.5).half().cuda()
res = (torch.randn((8, 768, 768, 128), generator=g) * 0.5).half().cuda()
w = torch.randn((128, 128, 3, 3), generator=g) * 0.03
b = torch.zeros(128)
for r in (None, res, None, res):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y = E.conv2d(x, w, b, mode=0, residual=r)
    torch.cuda.synchronize()
    print("residual" if r is not None else "plain   ", f"{(time.perf_counter() - t0) * 1e3:.2f} ms (incl. weight packing)")
