"""Uninitialised-read detector for the training step: the caching allocator's free memory is filled with NaN before the
step, so any kernel that reads bytes nobody wrote in this step turns its outputs into NaN (or visibly different numbers).
Prints which outputs differ from a clean run."""
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


def poison():
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    big = torch.full((int(os.environ.get("POISON_GB", "8")) << 28,), float("nan"), dtype=torch.float32, device=dev)
    small = [torch.full((n,), float("nan"), dtype=torch.float32, device=dev) for n in (64, 256, 1024, 4096, 16384, 65536) for _ in range(64)]
    torch.cuda.synchronize()
    del big, small


def grads(do_poison):
    m = lcasr_b200.SCConformerXL(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    if do_poison:
        poison()
    out = m(x)
    loss = ctc(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl)
    loss.backward()
    torch.cuda.synchronize()
    g = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    g["__loss__"] = loss.detach().reshape(1).clone()
    g["__logp__"] = out["final_posteriors"].detach().clone()
    return g


a = grads(False)
top = max(v.norm().item() for k, v in a.items() if not k.startswith("__"))
for trial in range(3):
    b = grads(True)
    bad = {}
    for n in a:
        d = (a[n] - b[n])
        nan = int(torch.isnan(b[n]).sum())
        rel = d[~torch.isnan(d)].norm().item() / max(a[n].norm().item(), 1e-3 * top) if d.numel() else 0.0
        if nan or rel > 1e-6:
            bad[n] = (nan, f"{rel:.1e}")
    print(f"trial {trial}: {len(bad)} of {len(a)} tensors differ from the clean run:", bad)
