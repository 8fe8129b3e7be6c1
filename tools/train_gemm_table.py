"""Per-shape timing table of every GEMM launch of one cfg-5 training step (CUDA events around each C-ABI call).
usage (GPU box): python tools/train_gemm_table.py [B] [T] > gpurun_out/train_gemm_table.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lcasr_b200
from lcasr_b200 import _lib as L
from oracle import lcasr_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
cfg = O.make_config(**O.BASELINE_MODELS["cfg5_6L768D6H"])
dev = torch.device("cuda", 0)
model = lcasr_b200.SCConformerXL(**cfg)
model.load_state_dict(O.synth_state_dict(cfg, seed=12345))
model = model.to(dev).train()
x = O.synth_input(B, T, 80, seed=1).to(dev)
N = O.calc_length(T)
tgt, tl = O.synth_targets(B, N, vocab=4095)
ctc = lcasr_b200.CTCLoss(blank=4095, reduction="sum")


def step():
    out = model(x)
    loss = ctc(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl)
    for p in model.parameters():
        p.grad = None
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
L.TIMING, L.TIMING_TAGS = {}, True
step()
torch.cuda.synchronize()
rec = L.timing_summary(L.TIMING)
L.TIMING, L.TIMING_TAGS = None, False
tot = sum(v[0] for v in rec.values())
print(f"instrumented step: {tot:.2f} ms in {sum(v[1] for v in rec.values())} launches")
for k, (ms, n) in sorted(rec.items(), key=lambda kv: -kv[1][0]):
    extra = ""
    if "[" in k and "x" in k:
        dims = k[k.index("[") + 1:].split(" ")[0].split("x")
        m, n_, kk = (int(v) for v in dims)
        batch = int(k.split("batch=")[1].rstrip("]")) if "batch=" in k else 1
        extra = f"  {2.0 * m * n_ * kk * batch * n / (ms / 1e3) / 1e12:7.1f} TFLOP/s"
    print(f"{ms:8.3f} ms  x{n:<3d} {ms / n * 1e3:8.1f} us/launch  {k}{extra}")
