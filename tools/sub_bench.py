#!/usr/bin/env python
"""time the fused subsampling level 1 (conv0 + SiLU + depthwise 3x3 s2): LCASR_SUB_TC=1 tensor-core conv0, 0 = SIMT kernel"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcasr_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
for B, T, C in ((1, 131072, 256), (16, 16384, 256), (1, 360000, 512)):
    spec = torch.randn(B, 80, T, generator=g).to(dev)
    w0, b0 = torch.randn(C, 9, generator=g).to(dev), (0.1 * torch.randn(C, generator=g)).to(dev)
    w1, b1 = (torch.randn(C, 9, generator=g) / 3).to(dev), (0.1 * torch.randn(C, generator=g)).to(dev)
    for _ in range(3):
        out = ops.subsample_conv0_dw(spec, w0, b0, w1, b1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = ops.subsample_conv0_dw(spec, w0, b0, w1, b1)
    e1.record(); torch.cuda.synchronize()
    print(f"SUB_TC={os.environ.get('LCASR_SUB_TC', 'default')} B={B} T={T} C={C}: {e0.elapsed_time(e1) / 5:.3f} ms  checksum {out.float().abs().mean().item():.6f}")
