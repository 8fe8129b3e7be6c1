"""Where does run-to-run variation of the training gradients start?  Repeats one step in one process with the backward's
intermediate tensors recorded (TrainEngine.trace) and reports, for every run that differs from the first, the FIRST
intermediate that differs and by how much."""
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
ctc = lcasr_b200.CTCLoss(blank=255, reduction="sum")
junk = torch.randn(32 << 20, device=dev)
side = torch.cuda.Stream(device=dev)


def run(perturb):
    m = lcasr_b200.SCConformerXL(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    m._train_engine = TrainEngine(m)
    m._train_engine.trace = []
    out = m(x)
    loss = ctc(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl)
    if perturb:  # unrelated kernels sharing the GPU change the order in which atomics land
        with torch.cuda.stream(side):
            for _ in range(200):
                junk.mul_(1.0000001)
    loss.backward()
    torch.cuda.synchronize()
    tr = m._train_engine.trace
    tr.append(("worst parameter gradient", torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None])))
    return tr


ref = run(False)
for i in range(int(os.environ.get("RUNS", "8"))):
    cur = run(i % 2 == 1)
    first = None
    for (name, a), (_, b) in zip(ref, cur):
        if not torch.equal(a, b):
            d = (a.float() - b.float())
            first = (name, f"{int((d != 0).sum())} of {d.numel()} elements differ", f"rel L2 {d.norm().item() / max(a.float().norm().item(), 1e-30):.2e}")
            break
    last = (ref[-1][1] - cur[-1][1]).norm().item() / ref[-1][1].norm().item()
    print(f"run {i} ({'perturbed' if i % 2 else 'alone'}): first differing intermediate: {first}; all parameter gradients rel L2 {last:.2e}")
