"""one training-form CTC launch (alpha and beta concurrently) at cfg-5 size for an ncu capture"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcasr_b200 import ops
from oracle import lcasr_oracle as O
dev = torch.device("cuda", 0)
B, N, V = 8, 2048, 4096
lp = torch.randn(B, N, V, device=dev).log_softmax(-1)
tgt, tl = O.synth_targets(B, N, vocab=V - 1)
il = torch.full((B,), N, dtype=torch.int32, device=dev)
for _ in range(2):
    ops.ctc_loss_fwd_ab(lp, tgt.to(dev), il, tl.to(dev), V - 1)
torch.cuda.synchronize()
