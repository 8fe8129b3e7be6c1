#!/usr/bin/env python
"""Per-rank COMPUTE cost of the native sequence-parallel forward without any communication: all ranks' phases run back to
back on ONE GPU (lcasr_model_forward_seqpar_emulated); time / world = what one rank computes.  Under
`ncu --metrics gpu__time_duration.sum` the launch list shows the kernels at a rank's shapes.
  python tools/sp_emulated_profile.py --world 8 [--model cfg3_6L768D24H --frames 131072] [--once]"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lcasr_b200
from lcasr_b200 import seqpar
from oracle import lcasr_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="cfg3_6L768D24H"); ap.add_argument("--frames", type=int, default=131072)
ap.add_argument("--world", type=int, default=8); ap.add_argument("--once", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = O.make_config(**O.BASELINE_MODELS[a.model])
model = lcasr_b200.SCConformerXL(**cfg, compute_dtype="bf16")
model.load_state_dict(O.synth_state_dict(cfg, seed=12345), strict=True)
model = model.to(dev).eval()
x = O.synth_input(1, a.frames, seed=1234).to(dev)
if a.once:
    seqpar.forward_sequence_parallel_emulated(model, x, a.world)
    torch.cuda.synchronize()
    sys.exit(0)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


t1 = timed(lambda: model(x))
te = timed(lambda: seqpar.forward_sequence_parallel_emulated(model, x, a.world))
print(json.dumps({"model": a.model, "frames": a.frames, "world": a.world, "ms_single_gpu": t1, "ms_all_ranks_emulated": te,
                  "ms_per_rank_compute": te / a.world, "compute_efficiency_bound": t1 / te}))
