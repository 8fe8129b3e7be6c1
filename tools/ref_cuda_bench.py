#!/usr/bin/env python
"""Head-to-head on ONE B200: the reference's own CUDA path and flash-attn alone beside our kernels.

  python tools/ref_cuda_bench.py [--skip-model]

(1) attention alone: flash_attn_func (the installed flash-attn 2.8.x wheel: HMMA/mma.sync cubins recompiled for sm_100 —
    what lcasr/components/attention.py:519-535 calls) vs lcasr_attention (tcgen05) at the two BASELINE shapes
    (H, Dh, N) = (24, 32, 16384) and (16, 128, 45000), bf16, CUDA events, 3 warm-up + 5 timed launches each.
(2) the UNMODIFIED reference model (baseline/_ref) on the GPU, eval(), flash_attn=True, under
    torch.autocast('cuda', bfloat16) exactly as exp/train.py:244 runs it, vs lcasr_b200.SCConformerXL on the same default-init
    weights and input: forward + greedy decode, cfg 2 (B = 4 to fit the reference's activations) and cfg 3.
Prints one JSON object."""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lcasr_b200
from lcasr_b200 import ops
from oracle import lcasr_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--skip-model", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
out = {"gpu": torch.cuda.get_device_name(0)}


def timed(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


try:
    from flash_attn import flash_attn_func
    import flash_attn
    out["flash_attn_version"] = flash_attn.__version__
except Exception as e:  # noqa: BLE001
    flash_attn_func = None
    out["flash_attn_error"] = f"{type(e).__name__}: {e}"

rows = []
for H, Dh, N in ((24, 32, 16384), (16, 128, 45000), (6, 128, 2048)):
    B = 16 if N == 2048 else 1
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(B, N, H, Dh, generator=g).bfloat16().to(dev) for _ in range(3))
    flops = 4.0 * B * H * N * N * Dh
    r = {"H": H, "Dh": Dh, "N": N, "B": B}
    ms = timed(lambda: ops.attention(q, k, v))
    r["lcasr_tcgen05_ms"], r["lcasr_tcgen05_tflops"] = ms, flops / ms / 1e9
    if flash_attn_func is not None:
        try:
            ms = timed(lambda: flash_attn_func(q, k, v))
            r["flash_attn_ms"], r["flash_attn_tflops"] = ms, flops / ms / 1e9
            r["speedup_over_flash_attn"] = r["flash_attn_ms"] / r["lcasr_tcgen05_ms"]
            r["max_abs_diff"] = (flash_attn_func(q, k, v).float().reshape(B, N, H * Dh) - ops.attention(q, k, v).float()).abs().max().item()
        except Exception as e:  # noqa: BLE001
            r["flash_attn_error"] = f"{type(e).__name__}: {e}"
    rows.append(r)
    del q, k, v
out["attention"] = rows

if not a.skip_model:
    from oracle.ref_import import load_reference, reference_available
    models = []
    for mkey, T, B in (("cfg2_9L768D6H", 16384, 4), ("cfg3_6L768D24H", 131072, 1)):
        cfg = O.make_config(**O.BASELINE_MODELS[mkey])
        r = {"model": mkey, "frames": T, "batch": B}
        x = O.synth_input(B, T, cfg["feat_in"], seed=1234).to(dev)
        torch.manual_seed(12345)
        ours = lcasr_b200.SCConformerXL(**cfg, compute_dtype="bf16")
        sd = {k_: v_.detach().clone() for k_, v_ in ours.state_dict().items()}
        ours = ours.to(dev).eval()

        def ours_step():
            ours(x)
            return ops.greedy_collapse(ours.last_argmax, cfg["vocab_size"])
        r["lcasr_b200_ms"] = timed(ours_step)
        lp_ours = ours(x)["final_posteriors"].float()
        try:
            assert reference_available(), "no reference tree (baseline/_ref)"
            Ref, Dec = load_reference()
            torch.manual_seed(12345)
            ref = Ref(**cfg)
            ref.load_state_dict(sd, strict=True)
            ref = ref.to(dev).eval()

            def ref_step():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    o = ref(x)
                return o["final_posteriors"].argmax(-1)
            r["reference_cuda_bf16_autocast_ms"] = timed(ref_step, n=3, warm=2)
            r["speedup_over_reference_cuda"] = r["reference_cuda_bf16_autocast_ms"] / r["lcasr_b200_ms"]
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                lp_ref = ref(x)["final_posteriors"].float()
            r["max_abs_ours_vs_reference_cuda_bf16"] = (lp_ours - lp_ref).abs().max().item()
            del ref
        except Exception as e:  # noqa: BLE001
            r["reference_error"] = f"{type(e).__name__}: {str(e)[:300]}"
        r["audio_s_per_s_lcasr_b200"] = B * T / 100 / (r["lcasr_b200_ms"] / 1e3)
        if "reference_cuda_bf16_autocast_ms" in r:
            r["audio_s_per_s_reference_cuda"] = B * T / 100 / (r["reference_cuda_bf16_autocast_ms"] / 1e3)
        models.append(r)
        del ours
        torch.cuda.empty_cache()
    out["model"] = models
print(json.dumps(out))
