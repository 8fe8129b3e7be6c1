"""Data-parallel training check on real GPUs (run under torchrun, one rank per GPU):
the gradients a rank holds after loss.backward() with the in-backward NCCL all-reduce must equal the mean over
ranks of the gradients each rank computes alone on its own batch.
usage: torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import lcasr_b200
from lcasr_b200.training import TrainEngine
from oracle import lcasr_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = O.make_config(n_layers=2, d_model=256, n_heads=2, head_dim=128, subsampling_conv_channels=64, vocab_size=255)
sd = O.synth_state_dict(cfg, seed=1)
x = O.synth_input(2, 1024, 80, seed=100 + rank).to(dev)
tgt, tl = O.synth_targets(2, O.calc_length(1024), vocab=255, seed=7 + rank)
ctc = lcasr_b200.CTCLoss(blank=255, reduction="sum")


def grads(dp, overlap=True):
    m = lcasr_b200.SCConformerXL(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    m._train_engine = TrainEngine(m)
    if dp:
        m._train_engine.dp_group = dist.group.WORLD
        m._train_engine.dp_overlap = overlap
    out = m(x)
    ctc(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl).backward()
    return {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}


solo, solo2, dp = grads(False), grads(False), grads(True)
dp_end = grads(True, overlap=False)  # one all-reduce after the backward: must agree with the overlapped form
ov = max((dp[n] - dp_end[n]).norm().item() / max(dp_end[n].norm().item(), 1e-30) for n in dp
         if dp_end[n].norm().item() > 1e-3 * max(v.norm().item() for v in dp_end.values()))
if rank == 0:
    print(f"overlapped vs end-of-backward all-reduce: worst relative L2 difference {ov:.3e}")
# run-to-run noise of ONE rank (fp32 atomics of the split-K weight gradients / bias column sums: leaves of the graph)
noise = max((solo[n] - solo2[n]).norm().item() / max(solo[n].norm().item(), 1e-30) for n in solo
            if solo[n].norm().item() > 1e-3 * max(v.norm().item() for v in solo.values()))
worst, means = ("", 0.0), {}
for n, g in solo.items():
    means[n] = g.clone()
    dist.all_reduce(means[n])
    means[n] /= world
# a bias in front of a batch norm has a mathematically zero gradient: both sides hold rounding noise there
floor = 1e-3 * max(m.norm().item() for m in means.values())
for n, mean in means.items():
    if mean.norm().item() < floor:  # noise-level gradient: only require that it stays noise-level
        assert (dp[n] - mean).norm().item() < floor, n
        continue
    rel = (dp[n] - mean).norm().item() / mean.norm().item()
    if rel > worst[1]:
        worst = (n, rel)
offenders = sorted(((dp[n] - means[n]).norm().item() / max(means[n].norm().item(), floor), n) for n in means)[-3:]
if rank == 0:
    print("worst parameter:", worst, "| top offenders:", [(n, f"{v:.1e}") for v, n in offenders])
worst = worst[1]
t = torch.tensor([worst], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dp_check world={world}: worst relative L2 difference between in-backward all-reduced gradients and the mean of "
          f"per-rank gradients = {t.item():.3e}; run-to-run noise of a single rank on this step = {noise:.3e}")
    assert t.item() < 1e-5  # the step is reproducible (fp64 / fixed-point cross-CTA sums): only leaf split-K atomics reorder
dist.destroy_process_group()
