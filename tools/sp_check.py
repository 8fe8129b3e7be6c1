#!/usr/bin/env python
"""torchrun check of the sequence-parallel paths on real GPUs: every rank compares its token block of the NATIVE driver
(csrc/seqpar.cu: NCCL P2P ring order + partial-attention merge, one C-ABI call) with the single-GPU forward of the whole
recording; rank 0 times the native driver, the Python-driven all-gather driver of round 1 and the single-GPU forward.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sp_check.py [--model cfg3_6L768D24H]"""
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
ap.add_argument("--skip-python-driver", action="store_true")
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
ref = model(x)["final_posteriors"][0]
ref_am = model.last_argmax[0].clone()
ncomm = seqpar.NativeComm()
lp, am_full, (s, e) = seqpar.forward_sequence_parallel_native(model, x, ncomm)
torch.cuda.synchronize()
err = (lp - ref[s:e]).abs().max().item()
agree = (am_full == ref_am).float().mean().item()
lp2, am2, _ = seqpar.forward_sequence_parallel_native(model, x, ncomm)
repeatable = bool(torch.equal(lp, lp2) and torch.equal(am_full, am2))


def timed(fn):
    for _ in range(2):
        fn()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / a.steps


t_native = timed(lambda: seqpar.forward_sequence_parallel_native(model, x, ncomm))
t_py = None
if not a.skip_python_driver:
    pcomm = seqpar.DistComm()
    t_py = timed(lambda: seqpar.forward_sequence_parallel(model, x, pcomm))
t_1 = timed(lambda: model(x))
errs = [None] * world
dist.all_gather_object(errs, (rank, err, agree, repeatable))
if rank == 0:
    print(json.dumps({"world": world, "model": a.model, "frames": a.frames, "dtype": a.dtype,
                      "per_rank_(rank,max_abs_vs_single_gpu,argmax_agree,bit_repeatable)": errs,
                      "ms_native": t_native * 1e3, "ms_python_allgather_driver": None if t_py is None else t_py * 1e3,
                      "ms_single_gpu": t_1 * 1e3, "speedup_native": t_1 / t_native, "efficiency_native": t_1 / t_native / world,
                      "audio_s_per_s_native": a.frames / 100 / t_native}))
ncomm.close()
dist.destroy_process_group()
