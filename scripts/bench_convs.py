"""Micro-benchmark of the tcgen05 implicit-GEMM kernel on the heavy conv shapes of SURVEY.md App. B."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpercept_b200 import engine as E  # noqa: E402

SHAPES = [  # (N, H, W, Cin, Cout, ks, mode)
    (8, 768, 768, 128, 128, 3, 0), (8, 384, 384, 256, 256, 3, 0), (8, 192, 192, 512, 512, 3, 0),
    (8, 96, 96, 512, 512, 3, 0), (8, 384, 384, 256, 256, 3, 3), (8, 192, 192, 512, 512, 3, 3),
    (8, 96, 96, 320, 320, 3, 0), (8, 48, 48, 640, 640, 3, 0), (8, 24, 24, 1280, 1280, 3, 0),
    (8, 12, 12, 1280, 1280, 3, 0), (8, 96, 96, 320, 2560, 1, 0), (8, 96, 96, 1280, 320, 1, 0),
    (8, 768, 768, 128, 128, 3, 2),
]
res = []
for s in SHAPES:
    us, fl = E.bench_conv(torch.float16, *s, iters=5)
    tf = fl / us / 1e6
    res.append({"shape": s, "usec": us, "tflops_algorithmic": tf})
    print(f"{s}: {us:9.1f} us  {tf:7.1f} TFLOP/s (algorithmic)")
json.dump(res, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "bench_convs.json"), "w"), indent=1)
