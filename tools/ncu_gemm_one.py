"""one forward GEMM launch for an `ncu --set full` capture: python tools/ncu_gemm_one.py M N K"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lcasr_b200 import ops, _lib as L

M, N, K = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda", 0)
a = torch.randn(M, K, device=dev).bfloat16()
w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, out=out, act=L.ACT_GELU_TANH)
torch.cuda.synchronize()
