"""Launches the dominant implicit-GEMM shapes once each (for `ncu --set full -k regex:igemm_kernel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpercept_b200 import engine as E  # noqa: E402

for s in [(8, 768, 768, 128, 128, 3, 0), (8, 384, 384, 256, 256, 3, 0), (8, 96, 96, 320, 2560, 1, 0),
          (8, 768, 768, 32, 128, 1, 0)]:          # the last one = the K-packed stem (encoder.conv_in as a 1x1 GEMM, K = 27 of 64)
    us, fl = E.bench_conv(torch.float16, *s, iters=1)
    print(s, us, fl / us / 1e6)
# launch 16: the same 128->128 convolution with GroupNorm+SiLU applied in its operand path (GP_GN_FUSE=1, igemm_patch.cu)
os.environ["GP_GN_FUSE"] = "1"
g = torch.Generator().manual_seed(0)
x = torch.randn((8, 768, 768, 128), generator=g).half().cuda()
w = torch.randn((128, 128, 3, 3), generator=g) * 0.03
E.gn_conv3x3(x, 32, torch.ones(128), torch.zeros(128), 1e-6, True, w)
torch.cuda.synchronize()
print("gn-fused conv done")
