"""Is a layer's gradient slice really final when the backward hands it to the all-reduce?  Single GPU: the per-layer
reduce hook is replaced by a snapshot of the slice; at the end of the backward every snapshot must equal the slice's final
content bit for bit (a later kernel adding into an already reduced slice would make overlapped and end-of-backward
data-parallel reductions disagree)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lcasr_b200
from lcasr_b200.training import TrainEngine
from oracle import lcasr_oracle as O

dev = torch.device("cuda", 0)
cfg = O.make_config(n_layers=2, d_model=256, n_heads=2, head_dim=128, subsampling_conv_channels=64, vocab_size=255)
sd = O.synth_state_dict(cfg, seed=1)
x = O.synth_input(2, 1024, 80, seed=100).to(dev)
tgt, tl = O.synth_targets(2, O.calc_length(1024), vocab=255, seed=7)
m = lcasr_b200.SCConformerXL(**cfg)
m.load_state_dict(sd, strict=True)
m = m.to(dev).train()
eng = TrainEngine(m)
m._train_engine = eng
snaps = []


class _H:
    def wait(self):
        pass


def snap(flat, prefix):
    keys = [k for k in eng._layout if (k.startswith(prefix) if prefix else not k.startswith("layers."))]
    lo = min(eng._layout[k][0] for k in keys)
    hi = max(eng._layout[k][0] + ((eng._layout[k][1] + 3) // 4) * 4 for k in keys)
    snaps.append((prefix, lo, hi, flat[lo:hi].clone()))
    return _H()


eng._reduce_slice = snap
eng.dp_group = object()
eng.dp_average = False
out = m(x)
lcasr_b200.CTCLoss(blank=255, reduction="sum")(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl).backward()
torch.cuda.synchronize()
flat = eng.last_flat_grad
bad = 0
for prefix, lo, hi, s in snaps:
    cur = flat[lo:hi]
    diff = (cur - s)
    n = int((diff != 0).sum())
    print(f"slice {prefix or 'non-layer'} [{lo},{hi}): {n} elements changed after the hand-off; max |change| {diff.abs().max().item():.3e}")
    if n:
        bad += 1
        for k, (o, cnt, shape) in eng._layout.items():
            if lo <= o < hi:
                dn = int((diff[o - lo:o - lo + cnt] != 0).sum())
                if dn:
                    print(f"    {k}: {dn} of {cnt} changed, rel {diff[o - lo:o - lo + cnt].norm().item() / max(s[o - lo:o - lo + cnt].norm().item(), 1e-30):.2e}")
print("FINAL" if bad == 0 else "NOT FINAL")
