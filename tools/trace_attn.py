#!/usr/bin/env python
"""Phase timeline of the v2 attention kernel's CTA (0,0,0): prints, per K/V tile, the clock64 stamps of
the softmax warps (wait S, S ready, loaded, max done, exp done, P published) and of the MMA issuer."""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lcasr_b200
from lcasr_b200 import ops, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=16384); ap.add_argument("--H", type=int, default=24)
ap.add_argument("--Dh", type=int, default=32); ap.add_argument("--B", type=int, default=1)
a = ap.parse_args()
lib = ctypes.CDLL(L.LIB_PATH)
dev = torch.device("cuda", 0)
q, k, v = (torch.randn(a.B, a.N, a.H, a.Dh).bfloat16().to(dev) for _ in range(3))
for _ in range(2):
    ops.attention(q, k, v, impl=L.ATTN_TCGEN05)
torch.cuda.synchronize()
lib.lcasr_debug_attn_trace(1)
ops.attention(q, k, v, impl=L.ATTN_TCGEN05)
torch.cuda.synchronize()
lib.lcasr_debug_attn_trace(0)
KT = 32
n = 2 * KT * 8 + 4 * KT * 4
buf = (ctypes.c_longlong * n)()
lib.lcasr_debug_attn_trace_read(buf, n)
t0 = min(x for x in buf[: 2 * KT * 8] if x > 0)
print("tile j | waitS  Srdy   ld     max    exp    pub   | mma_seen mma_done   (cycles since first stamp)")
for j in range(8, 20):
    for t in range(2):
        s = [buf[(t * KT + j) * 8 + kk] - t0 for kk in range(6)]
        m = [buf[2 * KT * 8 + (t * KT + j) * 4 + kk] - t0 for kk in range(2)]
        print(f"{'AB'[t]} {j:3d} | " + " ".join(f"{x:6d}" for x in s) + " | " + " ".join(f"{x:8d}" for x in m))
    print()
