"""Per-K-chunk timeline of the patch-resident kernel's CTA 0 (gp_debug_patch_trace): where a chunk's time goes between the
TMA load, the GroupNorm transform and the MMA issue.  usage: python scripts/patch_trace.py [Cin Cout W H N]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpercept_b200 import engine as E  # noqa: E402

Cin, Cout, W, H, N = (int(v) for v in (sys.argv[1:6] + ["128", "128", "768", "768", "8"][len(sys.argv) - 1:]))
L = E.lib()
L.gp_debug_patch_trace.argtypes = [ctypes.c_void_p]
L.gp_debug_patch_trace.restype = None
g = torch.Generator().manual_seed(0)
x = torch.randn((N, H, W, Cin), generator=g).half().cuda()
w = torch.randn((Cout, Cin, 3, 3), generator=g) * 0.02
gamma, beta = torch.ones(Cin), torch.zeros(Cin)
for label, silu in (("GroupNorm+SiLU fused", True),):
    E.gn_conv3x3(x, 32, gamma, beta, 1e-6, silu, w)        # warm
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    L.gp_debug_patch_trace(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    E.gn_conv3x3(x, 32, gamma, beta, 1e-6, silu, w)
    e1.record()
    torch.cuda.synchronize()
    L.gp_debug_patch_trace(None)
    t = buf.cpu().view(64, 8)
    print(f"== {label}: {Cin}->{Cout} @ {H}x{W} x{N}: whole call {e0.elapsed_time(e1) * 1000:.0f} us (includes host packing)")
    print("chunk | xform: wait->row0  row0->done | mma: wait->ready ready->issued | period (ready to ready)")
    for i in range(4, 28):
        a = t[i]
        per = int(t[i + 1][5] - a[5])
        print(f"{i:5d} | {int(a[1] - a[0]):10d} {int(a[2] - a[1]):11d} | {int(a[5] - a[4]):12d} {int(a[6] - a[5]):12d} | {per}")
