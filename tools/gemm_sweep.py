"""Time lcasr_gemm (forward tcgen05 GEMM) over the shapes of the path; env LCASR_GEMM_CG / LCASR_GEMM_DEBUG select
kernel variants.  usage: python tools/gemm_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lcasr_b200 import ops, _lib as L

dev = torch.device("cuda", 0)
SHAPES = [(16384, 3072, 768, "bf16"), (16384, 2304, 768, "bf16"), (16384, 4096, 768, "bf16"), (16384, 768, 3072, "f32r"),
          (16384, 768, 768, "f32r"), (32768, 768, 768, "f32r"), (32768, 768, 3072, "f32r"), (16384, 768, 4096, "f32r"), (32768, 3072, 768, "bf16"), (45056, 8192, 2048, "bf16"), (45056, 2048, 8192, "f32r")]
tag = f"CG={os.environ.get('LCASR_GEMM_CG', 'auto')} DEBUG={os.environ.get('LCASR_GEMM_DEBUG', '0')}"
for M, N, K, kind in SHAPES:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    resid = torch.randn(M, N, device=dev) if kind == "f32r" else None
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if kind == "f32r" else torch.bfloat16)
    f = lambda: ops.gemm(a, w, resid=resid, alpha=0.5, out=out, act=L.ACT_GELU_TANH if kind == "bf16" else L.ACT_NONE)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{tag}  {M}x{N}x{K} {kind}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
