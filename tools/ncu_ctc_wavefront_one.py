"""a few loss-only wavefront CTC launches at the 1-hour size (N = 45000, S = 13500) for an ncu capture"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcasr_b200 import ops
dev = torch.device("cuda", 0)
N, V = int(os.environ.get("CTC_N", "45000")), 4096
S = int(0.3 * N)
g = torch.Generator().manual_seed(0)
lp = torch.randn(1, N, V, generator=g).log_softmax(-1).to(dev)
tgt = torch.randint(0, V - 1, (1, S), generator=g).to(dev)
il = torch.full((1,), N, dtype=torch.int32, device=dev)
tl = torch.full((1,), S, dtype=torch.int64, device=dev)
for _ in range(3):
    nll, _ = ops.ctc_loss_fwd(lp, tgt, il, tl, V - 1)
torch.cuda.synchronize()
print("nll", nll.tolist())
