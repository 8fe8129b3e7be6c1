"""fused level-1 subsampling forward + backward at cfg-5 size (B=8, T=16384, C=256): timing with CUDA events, or a
target for an ncu capture (-k regex:subsample_l1_bwd)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcasr_b200 import ops, train_ops as T
dev = torch.device("cuda", 0)
B, F, Tn, C = 8, 80, 16384, 256
g = torch.Generator().manual_seed(0)
spec = torch.randn(B, F, Tn, generator=g).to(dev)
w0, b0 = (torch.randn(C, 9, generator=g) * 0.3).to(dev), (torch.randn(C, generator=g) * 0.1).to(dev)
w1, b1 = (torch.randn(C, 9, generator=g) * 0.3).to(dev), (torch.randn(C, generator=g) * 0.1).to(dev)
d1 = ops.subsample_conv0_dw(spec, w0, b0, w1, b1)
dd1 = torch.randn(d1.shape, generator=g).to(torch.bfloat16).to(dev)
grads = [torch.zeros(C, 9, device=dev), torch.zeros(C, device=dev), torch.zeros(C, 9, device=dev), torch.zeros(C, device=dev)]
reps = int(os.environ.get("REPS", "5"))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
T.subsample_l1_bwd_(spec, w0, b0, w1, dd1, *grads)
torch.cuda.synchronize()
ev[0].record()
for i in range(reps):
    T.subsample_l1_bwd_(spec, w0, b0, w1, dd1, *grads)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print(f"subsample_l1_bwd: min {min(ts) * 1e3:.1f} us  median {sorted(ts)[len(ts) // 2] * 1e3:.1f} us")
