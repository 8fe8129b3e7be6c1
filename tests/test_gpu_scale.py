"""GPU parity AT THE SIZES BASELINE.json NAMES (the headline numbers must be numbers of a verified output):
attention at N = 16384 (Dh 32, 128 key tiles, rotary to position 16383 upstream) and N = 45000 (Dh 128, ragged last
tile) against chunked fp32 softmax(QK^T)V, and whole-model runs at the cfg-2 and cfg-3 shapes against the CPU oracle."""
import math

import pytest
import torch

from gpu_util import build_model, margin_mask, report
from oracle import lcasr_oracle as O

pytestmark = pytest.mark.gpu


def _sdpa_fp32_chunked(q, k, v, chunk):
    """q,k,v [1,N,H,Dh] bf16 CUDA -> fp32 [1,N,H*Dh]; plain softmax(QK^T/sqrt(Dh))V in fp32, `chunk` queries at a time."""
    _, N, H, Dh = q.shape
    kf, vf = k[0].float().transpose(0, 1), v[0].float().transpose(0, 1)  # [H,N,Dh]
    out = torch.empty(N, H, Dh, dtype=torch.float32, device=q.device)
    for s in range(0, N, chunk):
        qs = q[0, s:s + chunk].float().transpose(0, 1)  # [H,c,Dh]
        p = torch.softmax(torch.bmm(qs, kf.transpose(1, 2)) / math.sqrt(Dh), dim=-1)
        out[s:s + chunk] = torch.bmm(p, vf).transpose(0, 1)
    return out.reshape(1, N, H * Dh)


@pytest.mark.parametrize("Dh,N,H,chunk", [(32, 16384, 24, 2048), (128, 45000, 16, 1000)])
def test_attention_at_baseline_sizes(cuda_device, Dh, N, H, chunk):
    from lcasr_b200 import ops, _lib as L
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=cuda_device).manual_seed(11)
    # rows with sharp maxima (lazy O rescale) and flat rows; keys with a slowly drifting mean so that late tiles raise the max
    q = (torch.randn(1, N, H, Dh, generator=g, device=cuda_device) * torch.linspace(0.2, 3.0, N, device=cuda_device)[None, :, None, None]).bfloat16()
    k = (torch.randn(1, N, H, Dh, generator=g, device=cuda_device) + torch.linspace(-0.5, 0.5, N, device=cuda_device)[None, :, None, None]).bfloat16()
    v = torch.randn(1, N, H, Dh, generator=g, device=cuda_device).bfloat16()
    got = ops.attention(q, k, v, impl=L.ATTN_TCGEN05).float()
    ref = _sdpa_fp32_chunked(q, k, v, chunk)
    err = (got - ref).abs().max().item()
    report(test="attn_tc_baseline_size", Dh=Dh, N=N, H=H, max_abs=err, ref_absmax=ref.abs().max().item())
    assert torch.isfinite(got).all()
    assert err < 2e-2, f"tcgen05 attention at N={N}: max-abs {err}"


def _default_init_model(cfg, device, mode):
    import lcasr_b200
    torch.manual_seed(12345)  # exp/train.py:363 — the constructor reproduces the reference's default init (tests/test_default_init.py)
    m = lcasr_b200.SCConformerXL(**cfg, compute_dtype=mode)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m.to(device).eval(), sd


def test_cfg3_shape_default_init_vs_oracle(cuda_device):
    """BASELINE config 3 (6L-768D-24H, 131072 frames -> N = 16384) with the reference's default init: bf16 mode within the
    stated 2e-2, fp32 mode within 1e-4 and identical greedy tokens, CTC loss within 1e-3 relative."""
    import lcasr_b200
    cfg = O.make_config(**O.BASELINE_MODELS["cfg3_6L768D24H"])
    T = 131072
    x = O.synth_input(1, T, cfg["feat_in"], seed=1234)
    model, sd = _default_init_model(cfg, cuda_device, "bf16")
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        ref, ref_len = O.encoder_forward(sd, cfg, x)
    V = cfg["vocab_size"]
    tgt, tl = O.synth_targets(1, ref.shape[1], vocab=V, frac=0.3, seed=99)
    ref_nll = torch.nn.functional.ctc_loss(ref.transpose(0, 1), tgt, ref_len.long(), tl, blank=V, reduction="sum").item()
    xd = x.to(cuda_device)
    for mode, bar in (("bf16", 2e-2), ("fp32", 1e-4)):
        if mode == "fp32":
            model, _ = _default_init_model(cfg, cuda_device, "fp32")
        out = model(xd)
        lp = out["final_posteriors"].cpu()
        err = (lp - ref).abs().max().item()
        nll = lcasr_b200.CTCLoss(blank=V, reduction="sum")(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl).item()
        rel = abs(nll - ref_nll) / abs(ref_nll)
        agree = (lp.argmax(-1) == ref.argmax(-1)).float().mean().item()
        report(test="model_cfg3_shape_" + mode, max_abs=err, bar=bar, ctc_rel=rel, argmax_agree=agree, ref_absmax=ref.abs().max().item())
        assert err < bar, f"cfg3 shape, {mode}: posteriors off by {err}"
        assert rel < 1e-3
        if mode == "fp32":
            assert O.greedy_decode(lp[0], V) == O.greedy_decode(ref[0], V)
        else:
            safe = margin_mask(ref, 4e-2)
            assert bool((lp.argmax(-1) == ref.argmax(-1))[safe].all())
        del out, lp


def test_cfg2_shape_synthetic_weights_vs_oracle(cuda_device):
    """BASELINE config 2 (9L-768D-6H, Dh 128, 16384 frames -> N = 2048), two recordings, unit-gain synthetic weights
    (every op away from identity: perturbed norm gains, biases, BatchRenorm statistics)."""
    cfg = O.make_config(**O.BASELINE_MODELS["cfg2_9L768D6H"])
    x = O.synth_input(2, 16384, cfg["feat_in"], seed=1234)
    model, cfg, sd = build_model(cfg, cuda_device, "bf16", seed=12345)
    with torch.no_grad():
        ref, _ = O.encoder_forward(sd, cfg, x)
    scale = max(1.0, ref.abs().max().item() / 8)
    lp = model(x.to(cuda_device))["final_posteriors"].cpu()
    err = (lp - ref).abs().max().item()
    safe = margin_mask(ref, 4e-2 * scale)
    agree = lp.argmax(-1) == ref.argmax(-1)
    report(test="model_cfg2_shape_bf16", max_abs=err, ref_absmax=ref.abs().max().item(), argmax_agree=agree.float().mean().item())
    assert err < 2e-2 * scale * 2.5  # unit-gain weights: logits 3x the default init's; the un-widened bar is the default-init tests'
    assert bool(agree[safe].all())
    model32, _, _ = build_model(cfg, cuda_device, "fp32", seed=12345)
    lp32 = model32(x.to(cuda_device))["final_posteriors"].cpu()
    err32 = (lp32 - ref).abs().max().item()
    report(test="model_cfg2_shape_fp32", max_abs=err32)
    assert err32 < 1e-4 * scale
    assert [O.greedy_decode(lp32[b], 4095) for b in range(2)] == [O.greedy_decode(ref[b], 4095) for b in range(2)]


@pytest.mark.parametrize("B,T,t,P", [(1, 32768, 2, 3), (1, 36000, 3, 2), (3, 24576, 1, 4), (1, 16384, 4, 2)])
def test_attention_tail_split_equals_one_launch(cuda_device, B, T, t, P):
    """The wave-quantisation fix of the dense attention launch (lcasr_model_set_attention_tail): the last query-tile pairs of the
    last recording computed as key-range partial results on side streams + exact merge must give the single-launch result
    up to the bf16 rounding of the attention output (the partial results are merged in fp32)."""
    cfg = O.make_config(**O.BASELINE_MODELS["cfg1_6L256D8H"])
    x = O.synth_input(B, T, cfg["feat_in"], seed=77).to(cuda_device)
    model, _ = _default_init_model(cfg, cuda_device, "bf16")
    with torch.no_grad():
        model.set_attention_tail(0, 0)
        ref = model(x)["final_posteriors"].float()
        model.set_attention_tail(t, P)
        got = model(x)["final_posteriors"].float()
        got2 = model(x)["final_posteriors"].float()
    err = (got - ref).abs().max().item()
    report(test="attention_tail_split", B=B, T=T, t=t, P=P, max_abs=err)
    assert torch.isfinite(got).all()
    assert torch.equal(got, got2), "tail split is not repeatable (missing stream dependency?)"
    assert err < 8e-3, f"tail split differs from the single launch by {err}"
    # rows far from the tail see the same attention launch arithmetic: most of the output is bit-identical in layer 1 only,
    # so the check here is the bound above plus agreement of the confident argmaxes
    mask = margin_mask(ref, 4e-2)
    assert (got.argmax(-1)[mask] == ref.argmax(-1)[mask]).all()
