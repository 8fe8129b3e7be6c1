#!/usr/bin/env python
"""bench.py — audio-seconds/second of the encoder + CTC hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A step = one pass of the hot path over one batch of synthetic recordings:
SCConformerXL.forward (8x conv subsampling, L conformer layers, self-conditioned CTC head, log-softmax
+ per-frame argmax) followed by the greedy CTC collapse.  Default workload = the configuration the
metric is quoted on, BASELINE.json configs[2]: 6L-768D-24H (head dim 32) at 20-min context
(131072 frames), one recording per GPU; it fits one GPU.  N > 1: one process per GPU (torchrun), each
rank transcribes its own recording — recordings are independent, no data-path collective ("weak").
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model key in oracle.BASELINE_MODELS, frames T, recordings per GPU)
    "cfg1": ("cfg1_6L256D8H", 1024, 1),
    "cfg2": ("cfg2_9L768D6H", 16384, 16),
    "cfg3": ("cfg3_6L768D24H", 131072, 1),
    "cfg4": ("cfg4_3L2048D16H", 360000, 1),
    # training step: forward (train mode) + CTC loss + backward (+ gradient all-reduce when N > 1), batch 8 per GPU
    "cfg5": ("cfg5_6L768D6H", 16384, 8),
}
CATS = ["subsample", "norm", "gemm", "attention", "rope", "convmod", "softmax"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


def flops_forward(cfg, T, N):
    """Algorithmic forward FLOPs per recording (SURVEY §8d formulas; 2*m*n*k, full N x N attention)."""
    d, C, L, V1 = cfg["d_model"], cfg["subsampling_conv_channels"], cfg["n_layers"], cfg["vocab_size"] + 1
    sub = 2 * 9 * C * (T // 2) * 40 + 2 * 9 * C * (T // 4) * 20 + 2 * C * C * (T // 4) * 20 + 2 * 9 * C * N * 10 + 2 * C * C * N * 10 + 2 * 10 * C * d * N
    layer = N * (46 * d * d + 18 * d) + 4 * N * N * d
    sc = (L - 1) * N * 4 * d * V1 if cfg["self_conditioning"] else 0
    return sub + L * layer + sc + N * 2 * d * V1


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


INIT_SEED = 12345  # exp/train.py:363: the reference seeds torch and builds the model with PyTorch's default initialisers


def default_init_state_dict(cfg):
    """The reference's default initialisation.  `lcasr_b200.SCConformerXL(**cfg)` consumes the torch random stream exactly
    like the reference constructor (pinned by tests/test_default_init.py against a SHA-256 of the reference's state_dict),
    so seed + constructor gives the reference's weights without the reference being present."""
    import torch
    import lcasr_b200
    torch.manual_seed(INIT_SEED)
    m = lcasr_b200.SCConformerXL(**cfg)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def load_cpu_reference(cfg, sd):
    """-> (kind, forward(x) -> (log-probs [B,N,V1], lengths), greedy(lp_row) -> token list).  kind 'reference': the UNMODIFIED
    reference package (pip-installed under baseline/_ref, or /root/reference in the build container) through its own public
    API — eval(), torch.no_grad(), fp32, one forward over the whole context, GreedyCTCDecoder (eval/run_eval_cpu.sh-style);
    kind 'port': the oracle restatement, only if the reference cannot be imported."""
    import torch
    from oracle import lcasr_oracle as O
    V = cfg["vocab_size"]
    try:
        from oracle.ref_import import load_reference, reference_available
        if reference_available():
            Ref, Dec = load_reference()
            torch.manual_seed(INIT_SEED)
            model = Ref(**cfg)
            model.load_state_dict(sd, strict=True)
            model.eval()
            dec = Dec(tokenizer=None, blank_id=V)

            def fwd(x):
                with torch.no_grad():
                    out = model(x)
                return out["final_posteriors"], out["length"]
            return "reference", fwd, (lambda row: dec(row))
    except Exception as e:  # noqa: BLE001 — fall back to the port and say so
        print(f"bench.py: reference import failed ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)

    def fwd_port(x):
        with torch.no_grad():
            return O.encoder_forward(sd, cfg, x)
    return "port", fwd_port, (lambda row: O.greedy_decode(row, V))


def cpu_reference_run(cfg, sd, B, T, steps, warmup, budget_s, seed=1234):
    """The reference's CPU path on all host cores.  Returns a dict: audio-s/s, seconds per step, steps timed, cores, kind and
    the LAST step's outputs (log-probs, greedy tokens) so that the caller can check the CUDA result against them."""
    import torch
    from oracle import lcasr_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = O.synth_input(B, T, cfg["feat_in"], seed=seed)
    kind, fwd, greedy = load_cpu_reference(cfg, sd)
    last = {}

    def step():
        lp, ln = fwd(x)
        last["lp"], last["len"] = lp, ln
        last["greedy"] = [greedy(lp[b]) for b in range(B)]

    t_start = time.perf_counter()
    did_warm = 0
    for _ in range(warmup):
        if did_warm >= 1 and time.perf_counter() - t_start > 0.2 * budget_s:
            break
        step(); did_warm += 1
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    per_step = sum(times) / len(times)
    return dict(value=B * T / 100.0 / per_step, per_step=per_step, steps=len(times), cores=cores, warm=did_warm, kind=kind,
                lp=last["lp"], length=last["len"], greedy=last["greedy"])


def parity_report(cfg, lp_cuda, lp_ref, greedy_cuda, greedy_ref, ctc_cuda, dtype):
    """CUDA output vs the CPU reference output on the same weights and input (north_star's three criteria)."""
    import torch
    V = cfg["vocab_size"]
    B, N, _ = lp_ref.shape
    err = (lp_cuda - lp_ref).abs().max().item()
    agree = (lp_cuda.argmax(-1) == lp_ref.argmax(-1))
    bar = 2e-2 if dtype == "bf16" else 1e-4
    top2 = lp_ref.topk(2, dim=-1).values
    safe = (top2[..., 0] - top2[..., 1]) > 2 * bar  # frames whose fp32 top-1 margin exceeds twice the posterior tolerance
    from oracle import lcasr_oracle as O
    tgt, tl = O.synth_targets(B, N, vocab=V, frac=0.3, seed=99)
    ref_nll = torch.nn.functional.ctc_loss(lp_ref.transpose(0, 1), tgt, torch.full((B,), N, dtype=torch.long), tl, blank=V,
                                           reduction="sum").item()  # exp/train.py:104,249: CTCLoss(blank=V, reduction='sum')
    return {"max_abs": err, "bar": bar, "within_bar": bool(err < bar), "ref_absmax": lp_ref.abs().max().item(),
            "greedy_equal": greedy_cuda == greedy_ref, "argmax_agree": agree.float().mean().item(),
            "argmax_agree_on_safe_margin_frames": bool(agree[safe].all()) if bool(safe.any()) else None,
            "safe_margin_frames": safe.float().mean().item(),
            "ctc_cuda": ctc_cuda, "ctc_ref": ref_nll, "ctc_rel": abs(ctc_cuda - ref_nll) / abs(ref_nll),
            "note": "CUDA " + dtype + " output vs the CPU reference's fp32 output, same default-init weights (seed 12345) and input; "
                    "greedy_equal compares full token lists, the safe-margin line only frames whose fp32 top-1/top-2 margin "
                    "exceeds 2x the posterior bar (an untrained model's posteriors are near-uniform)"}


def flops_train_step(cfg, T, N):
    """forward + backward: every GEMM-shaped forward FLOP costs 2 more in the backward (data + weight gradient;
    attention: dQ, dK, dV, dP and the recomputed P = 2.5x the forward)."""
    d, L = cfg["d_model"], cfg["n_layers"]
    fwd = flops_forward(cfg, T, N)
    attn = L * 4 * N * N * d
    return fwd + 2 * (fwd - attn) + 2.5 * attn


def cpu_reference_train(cfg, B, T, budget_s):
    """The reference's training step on the host cores: train-mode forward, CTC loss, autograd backward (fp32) — the
    unmodified reference package when it imports (exp/train.py:236-262 without the CUDA autocast), else the oracle port.
    A bounded sample: `Bs` recordings of the batch.  Returns (audio-s/s, seconds, Bs, cores, kind)."""
    import torch
    from oracle import lcasr_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = default_init_state_dict(cfg)
    Bs = 1
    x = O.synth_input(Bs, T, cfg["feat_in"], seed=1234)
    N = O.calc_length(T)
    V = cfg["vocab_size"]
    tgt, tl = O.synth_targets(Bs, N, vocab=V, frac=0.3, seed=99)
    kind = "port"
    try:
        from oracle.ref_import import load_reference, reference_available
        if reference_available():
            Ref, _ = load_reference()
            torch.manual_seed(INIT_SEED)
            model = Ref(**cfg)
            model.load_state_dict(sd, strict=True)
            model.train()
            t0 = time.perf_counter()
            out = model(audio_signal=x, length=None)
            loss = torch.nn.CTCLoss(blank=V, reduction="sum")(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl).sum()
            loss.backward()
            per = time.perf_counter() - t0
            kind = "reference"
    except Exception as e:  # noqa: BLE001
        print(f"bench.py: reference training step failed ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        kind = "port"
    if kind == "port":
        t0 = time.perf_counter()
        O.training_step(sd, cfg, x, tgt, tl)
        per = time.perf_counter() - t0
    return Bs * T / 100.0 / per, per, Bs, cores, kind


def seqpar_bench(mkey, T, dtype, steps, dev, rank, world, comm, dist):
    """ONE recording across `world` ranks (BASELINE configs 3 / 4: sequence-parallel ring attention): the native driver
    (lcasr_model_forward_seqpar: K/V blocks as NCCL P2P transfers in ring order overlapped with attention, neighbour halo
    exchange) against the single-GPU forward of the same recording on the same rank.  Device-timed, max over ranks."""
    import torch
    import lcasr_b200
    from lcasr_b200 import seqpar
    from oracle import lcasr_oracle as O
    cfg = O.make_config(**O.BASELINE_MODELS[mkey])
    sd = default_init_state_dict(cfg)
    model = lcasr_b200.SCConformerXL(**cfg, compute_dtype=dtype)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    x = O.synth_input(1, T, cfg["feat_in"], seed=1234).to(dev)  # every rank holds the recording, reads only its slice
    V = cfg["vocab_size"]

    def timed(fn, k):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sp_step():
        lp, am, blk = seqpar.forward_sequence_parallel_native(model, x, comm)
        return lp, am, blk, lcasr_b200.ops.greedy_collapse(am.view(1, -1), V)

    def single_step():
        out = model(x)
        return out, lcasr_b200.ops.greedy_collapse(model.last_argmax, V)

    ms_sp = timed(sp_step, steps)
    ms_1 = timed(single_step, steps)
    lp, am, (s0, e0_), _ = sp_step()
    out1, _ = single_step()
    err = (lp - out1["final_posteriors"][0, s0:e0_]).abs().max()
    agree = (am == model.last_argmax[0]).float().mean()
    stats = torch.stack([err.double(), (1.0 - agree).double()])
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    audio_s = T / 100.0
    del model
    torch.cuda.empty_cache()
    return {"model": mkey, "frames": T, "tokens": T // 8, "ranks": world, "ms_per_step": ms_sp, "ms_single_gpu": ms_1,
            "speedup_vs_single_gpu": ms_1 / ms_sp, "efficiency": ms_1 / ms_sp / world, "audio_s_per_s": audio_s / (ms_sp / 1e3),
            "max_abs_vs_single_gpu": float(stats[0].item()), "argmax_disagree_frac_vs_single_gpu": float(stats[1].item()),
            "what": "forward + log-softmax + greedy decode of ONE recording; tokens split into contiguous blocks; per layer the K/V "
                    "blocks of the other ranks arrive by one in-place NCCL all-gather on a side stream (LCASR_SP_KV_MODE=2; 0 = "
                    "ring-ordered ncclSend/ncclRecv steps) while the tcgen05 attention kernel already runs on the rank's own block, "
                    "then on the keys in front of and behind it in planned pieces on four streams (exact fp32 merge of the partial "
                    "results); conv-module halo rows to/from the two neighbours; one C-ABI call per rank and step"}


def dp_train_bench(steps, dev, rank, world, dist):
    """BASELINE config 5 under data parallelism: batch 8 per GPU, gradients all-reduced (averaged) per layer inside backward over
    NCCL; weak-scaling efficiency = the same rank's step WITHOUT the all-reduce / the data-parallel step."""
    import torch
    import lcasr_b200
    from lcasr_b200.training import TrainEngine
    from oracle import lcasr_oracle as O
    mkey, T, B = WORKLOADS["cfg5"]
    cfg = O.make_config(**O.BASELINE_MODELS[mkey])
    N, V = O.calc_length(T), cfg["vocab_size"]
    model = lcasr_b200.SCConformerXL(**cfg, compute_dtype="bf16")
    model.load_state_dict(default_init_state_dict(cfg), strict=True)
    model = model.to(dev).train()
    model._train_engine = TrainEngine(model)
    ctc = lcasr_b200.CTCLoss(blank=V, reduction="sum")
    x = O.synth_input(B, T, cfg["feat_in"], seed=1234 + rank).to(dev)
    tgt, tl = O.synth_targets(B, N, vocab=V, frac=0.3, seed=99 + rank)
    tgt, tl = tgt.to(dev), tl.to(dev)

    def step():
        out = model(audio_signal=x, length=None)
        loss = ctc(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl).sum()
        for p_ in model.parameters():
            p_.grad = None
        loss.backward()

    def timed(k):
        for _ in range(3):
            step()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            step()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    model._train_engine.dp_group = None
    ms_local = timed(steps)
    model._train_engine.dp_group = dist.group.WORLD
    ms_dp = timed(steps)
    del model
    torch.cuda.empty_cache()
    audio_s = B * T / 100.0
    return {"model": mkey, "frames": T, "batch_per_gpu": B, "ranks": world, "ms_per_step": ms_dp, "ms_per_step_without_allreduce": ms_local,
            "weak_scaling_efficiency": ms_local / ms_dp, "audio_s_per_s": world * audio_s / (ms_dp / 1e3),
            "what": "training step (train-mode forward + CTC loss + hand-written backward) with the per-layer asynchronous NCCL "
                    "all-reduce of the flat gradient buffer inside backward"}


def train_bench(args, cfg, mkey, T, B, rank, world, local):
    """cfg 5: the training step through the drop-in classes exactly as exp/train.py:236-262 drives them."""
    import torch
    import lcasr_b200
    from lcasr_b200 import _lib as L
    from oracle import lcasr_oracle as O
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner / warnings go to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N = O.calc_length(T)
    V = cfg["vocab_size"]
    sd = default_init_state_dict(cfg)
    model = lcasr_b200.SCConformerXL(**cfg, compute_dtype="bf16")
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).train()
    from lcasr_b200.training import TrainEngine
    model._train_engine = TrainEngine(model)
    if world > 1:
        model._train_engine.dp_group = dist.group.WORLD  # gradients all-reduced (averaged) inside backward
    ctc = lcasr_b200.CTCLoss(blank=V, reduction="sum")
    x_host = O.synth_input(B, T, cfg["feat_in"], seed=1234 + rank).pin_memory()
    tgt_host, tl_host = O.synth_targets(B, N, vocab=V, frac=0.3, seed=99 + rank)
    tgt_host, tl_host = tgt_host.pin_memory(), tl_host.pin_memory()
    x, tgt, tl = x_host.to(dev), tgt_host.to(dev), tl_host.to(dev)

    def step(xd, tg, tln):
        out = model(audio_signal=xd, length=None)
        loss = ctc(out["final_posteriors"].transpose(0, 1), tg, out["length"], tln).sum()
        for p in model.parameters():
            p.grad = None
        loss.backward()
        return loss

    for _ in range(max(args.warmup, 3)):
        step(x, tgt, tl)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    L.lib.lcasr_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(x, tgt, tl)
    e1.record()
    barrier()
    launches = int(L.lib.lcasr_launch_count())
    clocks = sampler.stop() if sampler else None
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    audio_s = B * T / 100.0
    value = world * audio_s / (ms_per_step / 1e3)

    # end to end: pinned host batch -> H2D -> step -> loss value back on the host
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(x_host.to(dev, non_blocking=True), tgt_host.to(dev, non_blocking=True), tl_host.to(dev, non_blocking=True))
        loss_host = loss.item()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps

    # beyond the metric (BASELINE config 5 ends at the gradient all-reduce): the device-side optimizer step of f2
    opt = lcasr_b200.optim.MADGRAD(model.parameters(), lr=1e-3)
    opt.max_grad_norm = 0.8
    step(x, tgt, tl); opt.step()  # state allocation
    torch.cuda.synchronize()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for _ in range(5):
        opt.step()
    o1.record()
    torch.cuda.synchronize()
    opt_ms = o0.elapsed_time(o1) / 5
    del opt

    # kernel breakdown: two instrumented steps (CUDA events around every C-ABI call on the launch stream)
    L.TIMING = {}
    for _ in range(2):
        step(x, tgt, tl)
    torch.cuda.synchronize()
    rec, L.TIMING = L.timing_summary(L.TIMING), None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks = load_peaks()
    kern = {k: {"ms_per_step": v[0] / 2, "launches_per_step": v[1] / 2} for k, v in sorted(rec.items(), key=lambda kv: -kv[1][0])}
    attn_bwd_ms = rec.get("lcasr_attention_bwd_flash", (0, 0))[0] / 2  # prep + dK/dV + dQ kernels (0 with the materialised form)
    gemm_ms = (rec.get("lcasr_gemm", (0, 0))[0] + rec.get("lcasr_gemm_ex", (0, 0))[0]) / 2 + attn_bwd_ms
    attn_fwd_ms = rec.get("lcasr_attention_train", (0, 0))[0] / 2
    total_flops = flops_train_step(cfg, T, N) * B
    attn_fwd_flops = cfg["n_layers"] * 4.0 * N * N * cfg["d_model"] * B
    gemm_flops = total_flops - attn_fwd_flops  # everything GEMM-shaped except the fused forward attention kernel
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    attn_bwd_flops = cfg["n_layers"] * 10.0 * N * N * cfg["d_model"] * B  # the 5 products of the definition (the kernels run 7)
    roofline = {"kernel": "tcgen05 GEMMs of the step (gemm_tc_kernel forward, gemm_tcx_kernel backward) + the flash-style "
                          "attention backward kernels, counted with the 5 tile products of the definition", "bound": "tensor", "achieved": achieved, "peak": peaks["tf_sust"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tf_sust"], "traffic": None,
                "peak_source": peaks["source"] + ", sustained bf16 figure (kernels timed inside a long step)",
                "launch_ms": gemm_ms, "algorithmic_flops_per_launch": gemm_flops, "share_of_step": gemm_ms / ms_per_step,
                "attention_fwd_tflops": attn_fwd_flops / (attn_fwd_ms / 1e3) / 1e12 if attn_fwd_ms > 0 else None,
                "attention_bwd_tflops": attn_bwd_flops / (attn_bwd_ms / 1e3) / 1e12 if attn_bwd_ms > 0 else None,
                "model_tflops_per_step": total_flops / 1e12,
                "note": "achieved = (algorithmic FLOPs of all GEMM launches of one step) / (their summed CUDA-event time)"}
    config = {"workload": f"cfg5: lcasr {mkey} random-init, training step (train-mode forward + CTC loss + backward"
                          f"{' + gradient all-reduce' if world > 1 else ''}), {T} frames ({T / 6000:.1f} min) chunks, batch {B} per GPU",
              "frames": T, "tokens": N, "recordings_per_gpu": B, "parallelism": f"data-parallel x{world}" if world > 1 else "single GPU",
              "l2": "working set (>8 GB of saved activations per step) exceeds the 126 MB L2; no explicit flush"}
    line = {"metric": "audio-sec/sec training step (fwd+bwd+CTC)", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": world * audio_s / e2e_s, "unit": "audio-s/s",
                    "h2d_bytes_per_step": int(x_host.numel() * 4 + tgt_host.numel() * 8 + tl_host.numel() * 8),
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_s * 1e3, "loss": loss_host},
            "gpu_launches": launches, "roofline": roofline, "kernels": kern,
            "optimizer_step_ms": {"value": opt_ms, "what": "clip_grad_norm_(0.8) + MADGRAD over all parameters as two multi-tensor "
                                                            "kernels (lcasr_b200.optim); NOT part of the timed step / metric"}}
    if not args.no_cpu_baseline and world == 1:
        v, per, Bs, cores, kind = cpu_reference_train(cfg, B, T, args.cpu_budget_s)
        line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                "sample": f"one training step (train-mode forward + CTC + autograd backward, fp32) of {Bs} of the {B} "
                                          f"recordings on the host CPU ({cores} threads), {per:.1f} s"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="recordings per GPU (0 = workload default)")
    ap.add_argument("--frames", type=int, default=0, help="override context length in 10 ms frames")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-gpu-extras", action="store_true", help="N > 1: skip the sequence-parallel and data-parallel-training objects")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    args = ap.parse_args()

    from oracle import lcasr_oracle as O  # cpu_baseline leg / --impl reference only (never the product path)
    rank, world, local = dist_env()
    mkey, T, B = WORKLOADS[args.workload]
    T = args.frames or T
    B = args.batch or B
    cfg = O.make_config(**O.BASELINE_MODELS[mkey])
    N = O.calc_length(T)
    audio_s = B * T / 100.0
    config = {"workload": f"{args.workload}: lcasr {mkey} random-init (PyTorch default init, seed 12345), {T} frames ({T / 6000:.1f} min) "
                          f"context, {B} recording(s)/GPU, forward + CTC log-softmax + greedy decode"
                          + (" + CTC loss (S = 0.3 N targets)" if args.workload == "cfg4" else ""),
              "frames": T, "tokens": N, "recordings_per_gpu": B, "parallelism": f"batch-parallel x{args.gpus} (independent recordings)",
              "l2": "working set (>1 GB of activations per step) exceeds the 126 MB L2; no explicit flush"}

    if args.workload == "cfg5":
        if args.impl == "reference":
            if rank != 0:
                return 0
            v, per, Bs, cores, kind = cpu_reference_train(cfg, B, T, args.cpu_budget_s)
            print(json.dumps({"impl": "reference", "metric": "audio-sec/sec training step (fwd+bwd+CTC)", "value": v,
                              "unit": "audio-s/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": per * 1e3,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": config, "gpu_launches": 0,
                              "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                               "sample": f"one training step of {Bs} recording(s) on the host CPU"},
                              "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return 0
        return train_bench(args, cfg, mkey, T, B, rank, world, local)

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(cfg, default_init_state_dict(cfg), B, T, args.steps, args.warmup, 2.0 * args.cpu_budget_s)
        what = ("the UNMODIFIED reference package (baseline/_ref) through its own API: SCConformerXL.eval() forward + GreedyCTCDecoder"
                if r["kind"] == "reference" else "the oracle port (the reference package could not be imported)")
        line = {"impl": "reference", "metric": "audio-sec/sec encoder+CTC", "value": r["value"], "unit": "audio-s/s", "n_gpus": args.gpus,
                "steps": r["steps"], "warmup": r["warm"], "ms_per_step": r["per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": f"{r['steps']} of {args.steps} requested steps timed, each one full {T}-frame forward + greedy "
                                           f"decode of {B} recording(s) on the host CPU (fp32, {r['cores']} threads) by {what}"},
                "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import lcasr_b200
    from lcasr_b200 import _lib as L
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner / warnings go to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sd = default_init_state_dict(cfg)                 # the reference's default init under seed 12345 (no checkpoints offline)
    model = lcasr_b200.SCConformerXL(**cfg, compute_dtype=args.dtype)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    model.cuda_graphs = args.workload == "cfg1"  # the 10-s config is launch-bound: one graph launch instead of ~135
    x_host = O.synth_input(B, T, cfg["feat_in"], seed=1234 + rank).pin_memory()
    x = x_host.to(dev)
    dec = lcasr_b200.GreedyCTCDecoder(None, blank_id=cfg["vocab_size"])

    with_ctc = args.workload == "cfg4"  # BASELINE config 4: "encoder forward + CTC loss" (exp/train.py:249)
    V = cfg["vocab_size"]
    ctc = lcasr_b200.CTCLoss(blank=V, reduction="sum")
    tgt_host, tl_host = O.synth_targets(B, N, vocab=V, frac=0.3, seed=99)
    tgt, tl = tgt_host.to(dev), tl_host.to(dev)
    in_len = torch.full((B,), N, dtype=torch.int32, device=dev)

    def step_device():
        out = model(x)
        toks = lcasr_b200.ops.greedy_collapse(model.last_argmax, V)
        if with_ctc:
            return toks, ctc(out["final_posteriors"].transpose(0, 1), tgt, in_len, tl)
        return toks, None

    for _ in range(max(args.warmup, 3)):
        step_device()
    # ---- device-resident timing: exactly K steps between barriers, CUDA events, max over ranks ----
    L.call("lcasr_model_set_timing", model._handle, 1)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    L.lib.lcasr_reset_launch_count()
    model.graph_launches_replayed = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = int(L.lib.lcasr_launch_count()) + model.graph_launches_replayed
    clocks = sampler.stop() if sampler else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms = (ctypes.c_float * len(CATS))()
    cnt = (ctypes.c_int32 * len(CATS))()
    L.call("lcasr_model_get_timing", model._handle, ms, cnt, len(CATS))
    L.call("lcasr_model_set_timing", model._handle, 0)
    ms_per_step = ms_total / args.steps
    value = world * audio_s / (ms_per_step / 1e3)

    # ---- end to end through the public API with HOST buffers (H2D + forward + collapse + D2H tokens) ----
    def step_e2e():
        toks = model.transcribe_host(x_host)  # H2D + forward + collapse + D2H of the tokens inside the C ABI call
        if with_ctc:  # the log-probs stay on the device (model._logp_buf); the loss value comes back to the host
            lp = model._logp_buf[: B * N * (V + 1)].view(B, N, V + 1)
            return toks, float(ctc(lp.transpose(0, 1), tgt, in_len, tl).item())
        return toks, None

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        toks, loss_host = step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    e2e_value = world * audio_s / e2e_s

    # ---- N > 1: the two paths that actually communicate (the headline above is N independent recordings) ----
    multi = {}
    if world > 1 and not args.no_multi_gpu_extras:
        from lcasr_b200 import seqpar as _sp
        del model
        torch.cuda.empty_cache()
        comm = _sp.NativeComm()
        k = max(2, min(args.steps, 5))
        multi["seqpar"] = {"cfg3": seqpar_bench("cfg3_6L768D24H", 131072, args.dtype, k, dev, rank, world, comm, dist),
                           "cfg4": seqpar_bench("cfg4_3L2048D16H", 360000, args.dtype, max(2, k // 2), dev, rank, world, comm, dist)}
        comm.close()
        multi["dp_train"] = dp_train_bench(k, dev, rank, world, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    d, H, Dh = cfg["d_model"], cfg["n_heads"], cfg["head_dim"]
    kern = {}
    for name, m_, c_ in zip(CATS, ms, cnt):
        kern[name] = {"ms_per_step": m_ / args.steps, "launches_per_step": c_ / args.steps}
    attn_launch_s = (ms[3] / max(1, cnt[3])) / 1e3
    attn_flops = 4.0 * B * N * N * d                       # per launch (one layer): QK^T + PV, full N x N
    gemm_flops = flops_forward(cfg, T, N) * B - cfg["n_layers"] * attn_flops
    achieved = attn_flops / attn_launch_s / 1e12 if attn_launch_s > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except (ValueError, OSError):
            traffic = None
    roofline = {"kernel": "attention (flash, non-causal, fused softmax)", "bound": "tensor", "achieved": achieved,
                "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sust"], "traffic": traffic,
                "peak_source": peaks["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                "launch_ms": attn_launch_s * 1e3, "algorithmic_flops_per_launch": attn_flops,
                "share_of_step": (ms[3] / args.steps) / ms_per_step,
                "gemm_tflops": gemm_flops / ((ms[2] / args.steps) / 1e3) / 1e12 if ms[2] > 0 else None,
                "model_tflops_per_step": flops_forward(cfg, T, N) * B / 1e12}
    dh = d // cfg["n_heads"]
    if dh <= 64:
        # one MUFU.EX2 per score (16 per clock and SM; packed f16x2 / bf16x2 exponentials are two scalar MUFU ops on sm_100a):
        # at small head dims the exponentials, not the tensor cores, bound the kernel
        mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        ceil_tf = 4.0 * dh * 16 * 148 * mhz * 1e6 / 1e12
        roofline["mufu_ceiling"] = {"tflops": ceil_tf, "frac": achieved / ceil_tf if ceil_tf > 0 else None,
                                    "what": f"4*Dh FLOPs per exponential, 16 ex2/clk/SM x 148 SMs at {mhz:.0f} MHz"}

    line = {"metric": "audio-sec/sec encoder+CTC", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                    "d2h_bytes_per_step": int(B * N * 4 + B * 4 + (4 if with_ctc else 0)), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches, "roofline": roofline, "kernels": kern}
    line.update(multi)

    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(cfg, sd, B, T, 1, 0, args.cpu_budget_s)
        line["cpu_baseline"] = {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": r["kind"],
                                "sample": f"one full {T}-frame forward + greedy decode of {B} recording(s) on the host CPU (fp32 torch "
                                          f"CPU ops, {r['cores']} threads), {r['per_step']:.1f} s, by "
                                          + ("the unmodified reference package (baseline/_ref)" if r["kind"] == "reference" else "the oracle port")}
        # the headline is a number of a VERIFIED output: compare what the timed path produced with the CPU reference's output
        out = model(x)
        lp_cuda = out["final_posteriors"].cpu()
        toks_d, cnt_d = lcasr_b200.ops.greedy_collapse(model.last_argmax, V)
        toks_c, cnt_c = toks_d.cpu(), cnt_d.cpu()
        greedy_cuda = [toks_c[b, : int(cnt_c[b])].tolist() for b in range(B)]
        ctc_cuda = float(ctc(out["final_posteriors"].transpose(0, 1), tgt, in_len, tl).item())
        line["parity"] = parity_report(cfg, lp_cuda, r["lp"].float(), greedy_cuda, r["greedy"], ctc_cuda, args.dtype)
        line["parity"]["against"] = r["kind"]
        if args.dtype == "bf16":  # north_star's fp32-mode criteria on the same weights and input (1e-4, identical greedy tokens)
            del out, lp_cuda
            m32 = lcasr_b200.SCConformerXL(**cfg, compute_dtype="fp32")
            m32.load_state_dict(sd, strict=True)
            m32 = m32.to(dev).eval()
            o32 = m32(x)
            t32, c32 = lcasr_b200.ops.greedy_collapse(m32.last_argmax, V)
            t32, c32 = t32.cpu(), c32.cpu()
            g32 = [t32[b, : int(c32[b])].tolist() for b in range(B)]
            ctc32 = float(ctc(o32["final_posteriors"].transpose(0, 1), tgt, in_len, tl).item())
            p32 = parity_report(cfg, o32["final_posteriors"].cpu(), r["lp"].float(), g32, r["greedy"], ctc32, "fp32")
            line["parity"]["fp32_mode"] = {k: p32[k] for k in ("max_abs", "bar", "within_bar", "greedy_equal", "argmax_agree", "ctc_rel")}
            del m32, o32
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
