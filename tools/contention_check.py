"""Does the training step give the same gradients when other kernels share the GPU?  (A data-parallel run overlaps NCCL
all-reduces with the backward.)  Single GPU: gradients of one step alone vs the same step while a side stream keeps a
stream of unrelated kernels running."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lcasr_b200
from oracle import lcasr_oracle as O

dev = torch.device("cuda", 0)
cfg = O.make_config(n_layers=2, d_model=256, n_heads=2, head_dim=128, subsampling_conv_channels=64, vocab_size=255)
sd = O.synth_state_dict(cfg, seed=1)
x = O.synth_input(2, 1024, 80, seed=100).to(dev)
tgt, tl = O.synth_targets(2, O.calc_length(1024), vocab=255, seed=7)
ctc = lcasr_b200.CTCLoss(blank=255, reduction="sum")
side = torch.cuda.Stream(device=dev)
junk = torch.randn(64 << 20, device=dev)


def grads(noise):
    m = lcasr_b200.SCConformerXL(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    out = m(x)
    loss = ctc(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl)
    torch.cuda.synchronize()
    if noise:
        with torch.cuda.stream(side):
            for _ in range(400):       # memory-bound kernels with many CTAs, like an all-reduce
                junk.mul_(1.0000001)
    loss.backward()
    torch.cuda.synchronize()
    g = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    g["__loss__"] = loss.detach().reshape(1).clone()
    g["__logp__"] = out["final_posteriors"].detach().clone()
    return g


a, b, c = grads(False), grads(False), grads(True)
top = max(v.norm().item() for v in a.values())
for tag, u, v in (("alone vs alone", a, b), ("alone vs contended", a, c)):
    rows = sorted(((u[n] - v[n]).norm().item() / max(u[n].norm().item(), 1e-3 * top), n) for n in u)[-4:]
    print(tag, [(n, f"{r:.1e}") for r, n in rows])
    allr = {n: (u[n] - v[n]).norm().item() / max(u[n].norm().item(), 1e-3 * top) for n in u}
    print("   differing tensors:", sum(1 for r in allr.values() if r > 1e-6), "of", len(allr), "| loss", u["__loss__"].item(), v["__loss__"].item(),
          "| logp max-abs diff", (u["__logp__"] - v["__logp__"]).abs().max().item())
    print("   ", {n: f"{r:.1e}" for n, r in allr.items() if r > 1e-6})
