#!/usr/bin/env python
"""torchrun check of the sequence-parallel path on real GPUs: every rank compares its token block with
the single-GPU forward of the whole recording, and rank 0 times both.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sp_check.py [--workload cfg3]"""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lcasr_b200
from lcasr_b200 import seqpar
from oracle import lcasr_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="cfg3_6L768D24H"); ap.add_argument("--frames", type=int, default=131072)
ap.add_argument("--dtype", default="bf16"); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
cfg = O.make_config(**O.BASELINE_MODELS[a.model])
model = lcasr_b200.SCConformerXL(**cfg, compute_dtype=a.dtype)
model.load_state_dict(O.synth_state_dict(cfg, seed=12345), strict=True)
model = model.to(dev).eval()
x = O.synth_input(1, a.frames, seed=1234).to(dev)
comm = seqpar.DistComm()
ref = model(x)["final_posteriors"][0]
ref_am = model.last_argmax[0]
(part,), am_full = seqpar.forward_sequence_parallel(model, x, comm)
lp, am, (s, e) = part
err = (lp - ref[s:e]).abs().max().item()
agree = (am_full == ref_am).float().mean().item()
def timed(fn):
    for _ in range(2): fn()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.steps): fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / a.steps
t_sp = timed(lambda: seqpar.forward_sequence_parallel(model, x, comm))
t_1 = timed(lambda: model(x))
errs = [None] * world
dist.all_gather_object(errs, (rank, err, agree))
if rank == 0:
    print(json.dumps({"world": world, "model": a.model, "frames": a.frames, "dtype": a.dtype, "per_rank_max_abs_vs_single_gpu": errs,
                      "ms_sequence_parallel": t_sp * 1e3, "ms_single_gpu": t_1 * 1e3, "speedup": t_1 / t_sp,
                      "audio_s_per_s_sp": a.frames / 100 / t_sp}))
dist.destroy_process_group()
