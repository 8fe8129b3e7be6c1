"""Long-form moving-window inference throughput: one synthetic recording, windows of `seq_len` frames with
87.5 % overlap (the published evaluations' setting), cfg-2-style model.  Compares the batched device path
(lcasr_b200.longform) with the reference's control flow (one window per forward, posteriors copied to the host, exp /
count accumulation in torch on the CPU) driving the SAME CUDA model.
usage: python tools/longform_bench.py [minutes=20] [seq_len=4096] [max_batch=32]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lcasr_b200
from lcasr_b200.longform import transcribe_longform, plan_windows
from oracle import lcasr_oracle as O

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
seq_len = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
max_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
overlap = seq_len * 7 // 8
T = int(minutes * 6000)
dev = torch.device("cuda", 0)
cfg = O.make_config(**O.BASELINE_MODELS["cfg2_9L768D6H"])
model = lcasr_b200.SCConformerXL(**cfg)
model.load_state_dict(O.synth_state_dict(cfg, seed=12345))
model = model.to(dev).eval()
model.device = dev
spec = O.synth_input(1, T, 80, seed=9).pin_memory()
V = cfg["vocab_size"]


def reference_style():
    """lcasr/eval/utils.py:45-111 control flow on the CUDA model (what eval/run.py does today)"""
    wins, sl, ov = plan_windows(T, seq_len, overlap, 8)
    n_total = T // 4 + sl
    all_logits = torch.zeros(n_total, V + 1)
    count = torch.zeros(n_total, 1)
    position = 0
    for i, u in wins:
        out = model(spec[:, :, i:i + u].to(dev))
        probs = torch.exp(out["final_posteriors"][0].cpu())
        n = probs.shape[0]
        if i != 0:
            position -= int(ov / (u / n))
        all_logits[position:position + n] += probs
        count[position:position + n] += 1
        position += n
    lp = torch.log(all_logits[:position] / count[:position])
    return lcasr_b200.GreedyCTCDecoder(None, blank_id=V)(lp.to(dev))


for name, fn in (("batched_device", lambda: transcribe_longform(model, spec.to(dev), seq_len, overlap, max_batch=max_batch)),
                 ("reference_control_flow", reference_style)):
    toks = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    toks = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"path": name, "audio_minutes": minutes, "seq_len": seq_len, "overlap": overlap,
                      "windows": len(plan_windows(T, seq_len, overlap, 8)[0]), "seconds": dt, "audio_s_per_s": T / 100.0 / dt,
                      "tokens": len(toks)}))
