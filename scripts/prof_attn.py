"""One fused-attention launch at the UNet's 96x96 level (T=9216, 5 heads, d=64, B=8) for ncu, plus timing."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpercept_b200 import engine as E  # noqa: E402

B, T, H, D = 8, int(os.environ.get("T", 9216)), 5, 64
g = torch.Generator().manual_seed(0)
q, k, v = (torch.randn((B, T, H * D), generator=g).half().cuda() for _ in range(3))
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = E.attention(q, k, v, H, D ** -0.5)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"attention call {i}: {dt*1e3:.2f} ms (includes scale/vT GEMMs + setup), algorithmic {4.0*B*H*T*T*D/1e12:.3f} TF")
