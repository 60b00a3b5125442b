"""Debug: phase timeline of the fused attention kernel (CTA 0, tile A, row 0) from clock64() stamps."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpercept_b200 import engine as E  # noqa: E402

B, T, H, D = 8, 9216, 5, 64
g = torch.Generator().manual_seed(0)
q, k, v = (torch.randn((B, T, H * D), generator=g).half().cuda() for _ in range(3))
E.attention(q, k, v, H, D ** -0.5)
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
L = E.lib()
L.gp_debug_fattn_trace.argtypes = [ctypes.c_void_p]
L.gp_debug_fattn_trace.restype = None
L.gp_debug_fattn_trace(ctypes.c_void_p(buf.data_ptr()))
E.attention(q, k, v, H, D ** -0.5)
torch.cuda.synchronize()
t = buf.cpu().tolist()
L.gp_debug_fattn_trace(None)
print("softmax thread (tile A row 0): per-block phase durations in cycles")
print("blk  waitS ld+max waitPV rescale+turn exp+P fence | period")
for j in range(2, 26):
    s = t[j * 8: j * 8 + 7]
    nxt = t[(j + 1) * 8]
    print(f"{j:3d} {s[1]-s[0]:6d} {s[2]-s[1]:5d} {s[3]-s[2]:5d} {s[4]-s[3]:5d} {s[5]-s[4]:6d} {s[6]-s[5]:5d} | {nxt-s[0]:6d}")
print("MMA issuer (tile A), relative to the softmax thread's stamps of the same block j:")
print("blk  S(j) issued -> S(j) seen by softmax | P(j) arrive -> PV(j) issue start | PV issue cycles | S(j+1) issued - S(j) loaded")
for j in range(2, 12):
    m = t[512 + j * 4: 512 + j * 4 + 3]
    m1 = t[512 + (j + 1) * 4]
    s = t[j * 8: j * 8 + 7]
    print(f"{j:3d} {s[1]-m[0]:8d} {m[1]-s[6]:8d} {m[2]-m[1]:8d} {m1-s[1]:8d}")
