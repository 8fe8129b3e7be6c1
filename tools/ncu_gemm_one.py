"""one forward GEMM launch for an `ncu --set full` capture: python tools/ncu_gemm_one.py M N K"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lcasr_b200 import ops, _lib as L

M, N, K = (int(v) for v in sys.argv[1:4])
RESID = len(sys.argv) > 4 and sys.argv[4] == "f32r"   # fp32 output with a residual (out = resid + 0.5 * A W^T)
dev = torch.device("cuda", 0)
a = torch.randn(M, K, device=dev).bfloat16()
w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.float32 if RESID else torch.bfloat16)
resid = torch.randn(M, N, device=dev) if RESID else None
for _ in range(3):
    if RESID:
        ops.gemm(a, w, resid=resid, alpha=0.5, out=out)
    else:
        ops.gemm(a, w, out=out, act=L.ACT_GELU_TANH)
torch.cuda.synchronize()
