"""GPU parity of every unit operator of the C ABI against the oracle / torch fp32 on the same
seeded inputs.  fp32 kernels: ~1e-5 relative; bf16 kernels: <= ~1 bf16 ulp of the output scale."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import report
from oracle import lcasr_oracle as O

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize("d", [64, 256, 768, 2048, 200])
@pytest.mark.parametrize("kind", ["layer_norm", "rms_norm"])
def test_layernorm(cuda_device, d, kind):
    from lcasr_b200 import ops
    M = 333
    x = _rand(M, d, seed=1, scale=3.0) + 0.7
    w, b = 1 + 0.1 * _rand(d, seed=2), 0.1 * _rand(d, seed=3)
    if kind == "layer_norm":
        ref = F.layer_norm(x, (d,), w, b, 1e-5)
        eps = 1e-5
    else:
        ref = w * (x / (x.norm(2, dim=-1, keepdim=True) * d ** -0.5 + 1e-8))
        b, eps = None, 1e-8
    xc = x.to(cuda_device)
    o32, obf = ops.layernorm(xc, w.to(cuda_device), None if b is None else b.to(cuda_device), eps, kind, out_f32=True,
                             lo_dtype=torch.bfloat16)
    assert (o32.cpu() - ref).abs().max() < 2e-5
    assert (obf.float().cpu() - ref).abs().max() < 2 ** -8 * ref.abs().max()
    # in place (out_f32 aliases x)
    from lcasr_b200 import _lib as L
    ops_out, wc, bc = xc.clone(), w.to(cuda_device), None if b is None else b.to(cuda_device)
    L.call("lcasr_layernorm", ops_out.data_ptr(), wc.data_ptr(), L.ptr(bc), M, d, eps, 1 if kind == "rms_norm" else 0,
           ops_out.data_ptr(), None, 0, L.current_stream())
    assert (ops_out.cpu() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("C,T,B", [(32, 264, 2), (256, 1024, 1), (512, 77, 1)])
def test_subsampling_stencils(cuda_device, C, T, B):
    from lcasr_b200 import ops
    Fq = 80
    spec = _rand(B, Fq, T, seed=4)
    w0, b0 = _rand(C, 1, 3, 3, seed=5), 0.1 * _rand(C, seed=6)
    ref0 = F.silu(F.conv2d(spec.transpose(1, 2).unsqueeze(1), w0, b0, stride=2, padding=1))  # [B,C,T1,F1]
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2 ** -8)):
        got = ops.subsample_conv0(spec.to(cuda_device), w0.reshape(C, 9).contiguous().to(cuda_device), b0.to(cuda_device), dt)
        got = got.float().cpu().permute(0, 3, 1, 2)
        assert got.shape == ref0.shape
        assert (got - ref0).abs().max() < tol * max(1.0, ref0.abs().max())
    w1, b1 = _rand(C, 1, 3, 3, seed=7), 0.1 * _rand(C, seed=8)
    ref1 = F.conv2d(ref0, w1, b1, stride=2, padding=1, groups=C)
    x_cl = ref0.permute(0, 2, 3, 1).contiguous()
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2 ** -7)):
        got = ops.subsample_dwconv(x_cl.to(cuda_device, dt), w1.reshape(C, 9).contiguous().to(cuda_device), b1.to(cuda_device))
        got = got.float().cpu().permute(0, 3, 1, 2)
        assert got.shape == ref1.shape
        assert (got - ref1).abs().max() < tol * max(1.0, ref1.abs().max())


@pytest.mark.parametrize("M,N,K", [(66, 192, 64), (333, 128, 320), (1000, 4096, 256), (128, 256, 4096), (257, 72, 36)])
def test_gemm_simt_fp32(cuda_device, M, N, K):
    from lcasr_b200 import ops, _lib as L
    a, w = _rand(M, K, seed=9), _rand(N, K, seed=10) / math.sqrt(K)
    bias, resid = _rand(N, seed=11), _rand(M, N, seed=12)
    ref = a.double() @ w.double().T
    ac, wc = a.to(cuda_device), w.to(cuda_device)
    got = ops.gemm(ac, wc, impl=L.GEMM_SIMT)
    assert (got.cpu().double() - ref).abs().max() < 1e-5 * max(1.0, ref.abs().max())
    got = ops.gemm(ac, wc, bias=bias.to(cuda_device), act=L.ACT_GELU_TANH, impl=L.GEMM_SIMT)
    assert (got.cpu() - F.gelu((ref + bias).float(), approximate="tanh")).abs().max() < 2e-5
    got = ops.gemm(ac, wc, bias=bias.to(cuda_device), act=L.ACT_SILU, impl=L.GEMM_SIMT)
    assert (got.cpu() - F.silu((ref + bias).float())).abs().max() < 2e-5
    r = resid.to(cuda_device)
    got = ops.gemm(ac, wc, bias=bias.to(cuda_device), resid=r, alpha=0.5, impl=L.GEMM_SIMT, out=r)  # in place
    assert (got.cpu() - (resid + 0.5 * (ref + bias).float())).abs().max() < 2e-5


def test_glu_and_cast(cuda_device):
    from lcasr_b200 import ops
    x = _rand(77, 2 * 256, seed=13)
    ref = F.glu(x, dim=-1)
    assert (ops.glu(x.to(cuda_device)).cpu() - ref).abs().max() < 1e-6
    assert (ops.glu(x.to(cuda_device, torch.bfloat16)).float().cpu() - F.glu(x.bfloat16().float(), -1)).abs().max() < 2 ** -8 * 4


@pytest.mark.parametrize("Dh,base", [(32, 1500000), (128, 1500000), (64, 10000)])
def test_rope_table_and_split(cuda_device, Dh, base):
    from lcasr_b200 import ops
    N, B, H = 300, 2, 3
    inv_freq = 1.0 / (base ** (torch.arange(0, Dh, 2).float() / Dh))
    sd = {"rotary_pos_emb.inv_freq": inv_freq, "rotary_pos_emb.rotary_interpolation_factor": torch.tensor(1.0)}
    for off in (0, 44000):  # positions reach 45000 at 1-hour context
        cos, sin = ops.rope_table(inv_freq.to(cuda_device), 1.0, N, offset=off)
        # the fp32 angle the reference forms (rotary_emb.py:52-53), then cos/sin evaluated in float64:
        # torch's vectorised CPU cosf itself is only ~1e-4 accurate at |angle| ~ 4e4
        ang = (torch.arange(off, off + N).float()[:, None] * inv_freq[None, :]).double()
        assert (cos.cpu().double() - ang.cos()).abs().max() < 5e-7
        assert (sin.cpu().double() - ang.sin()).abs().max() < 5e-7
    cos_ref, sin_ref = O.rotary_tables(sd, N)
    cos, sin = ops.rope_table(inv_freq.to(cuda_device), 1.0, N)
    assert (cos.cpu() - cos_ref[:, : Dh // 2]).abs().max() < 2e-6
    assert (sin.cpu() - sin_ref[:, : Dh // 2]).abs().max() < 2e-6
    cos_ref, sin_ref = O.rotary_tables(sd, N)
    cos, sin = ops.rope_table(inv_freq.to(cuda_device), 1.0, N)
    d = H * Dh
    qkv = _rand(B * N, 3 * d, seed=14)
    q_ref, k_ref, v_ref = [t.reshape(B, N, H, Dh) for t in qkv.split(d, dim=-1)]
    c, s = cos_ref[None, :, None, :], sin_ref[None, :, None, :]
    q_rot = q_ref * c + O.rotate_half(q_ref) * s
    k_rot = k_ref * c + O.rotate_half(k_ref) * s
    q, k, v = ops.rope_split(qkv.to(cuda_device), B, N, H, Dh, cos, sin)
    assert (q.cpu() - q_rot).abs().max() < 1e-5 and (k.cpu() - k_rot).abs().max() < 1e-5
    assert torch.equal(v.cpu(), v_ref)
    q, k, vt = ops.rope_split(qkv.to(cuda_device), B, N, H, Dh, cos, sin, v_transposed=True)
    assert torch.equal(vt.cpu()[..., :N], v_ref.permute(0, 2, 3, 1))
    q, k, v = ops.rope_split(qkv.to(cuda_device), B, N, H, Dh, None, None)
    assert torch.equal(q.cpu(), q_ref) and torch.equal(k.cpu(), k_ref)


@pytest.mark.parametrize("Dh,N,H,B", [(32, 33, 2, 2), (128, 125, 1, 1), (32, 300, 3, 1), (64, 64, 2, 1), (128, 257, 2, 2)])
def test_attention_simt(cuda_device, Dh, N, H, B):
    from lcasr_b200 import ops, _lib as L
    q, k, v = (_rand(B, N, H, Dh, seed=s) for s in (15, 16, 17))
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).reshape(B, N, H * Dh)
    got = ops.attention(q.to(cuda_device), k.to(cuda_device), v.to(cuda_device), impl=L.ATTN_SIMT)
    assert (got.cpu() - ref).abs().max() < 2e-5
    gb = ops.attention(q.to(cuda_device, torch.bfloat16), k.to(cuda_device, torch.bfloat16), v.to(cuda_device, torch.bfloat16),
                       impl=L.ATTN_SIMT)
    assert (gb.float().cpu() - ref).abs().max() < 3e-2


@pytest.mark.parametrize("d,N,B", [(64, 50, 2), (256, 128, 1), (768, 333, 1)])
def test_dwconv_brn_silu(cuda_device, d, N, B):
    from lcasr_b200 import ops
    x = _rand(B, N, d, seed=18)
    w, b = _rand(d, 1, 9, seed=19) / 3, 0.1 * _rand(d, seed=20)
    rm, rs = 0.1 * _rand(d, seed=21), 0.75 + 0.5 * torch.rand(d, generator=torch.Generator().manual_seed(22))
    bw, bb = 1 + 0.1 * _rand(d, seed=23), 0.1 * _rand(d, seed=24)
    y = F.conv1d(x.transpose(1, 2), w, b, padding=4, groups=d)
    y = (y - rm[None, :, None]) / rs[None, :, None] * bw[None, :, None] + bb[None, :, None]
    ref = F.silu(y).transpose(1, 2)
    args = [t.to(cuda_device) for t in (w.reshape(d, 9).contiguous(), b, rm, rs, bw, bb)]
    got = ops.dwconv_brn_silu(x.to(cuda_device), *args)
    assert (got.cpu() - ref).abs().max() < 1e-5
    gb = ops.dwconv_brn_silu(x.to(cuda_device, torch.bfloat16), *args)
    assert (gb.float().cpu() - ref).abs().max() < 2 ** -7 * max(1.0, ref.abs().max())


@pytest.mark.parametrize("V", [128, 256, 4096])
def test_softmax_logsoftmax_argmax(cuda_device, V):
    from lcasr_b200 import ops
    x = _rand(77, V, seed=25, scale=3.0)
    x[5, 17] = x[5].max() + 1.0
    x[5, 90] = x[5, 17]  # tie: first index must win (torch.argmax semantics)
    assert (ops.softmax(x.to(cuda_device)).cpu() - x.softmax(-1)).abs().max() < 1e-6
    sb = ops.softmax(x.to(cuda_device, torch.bfloat16)).float().cpu()
    assert (sb - x.bfloat16().float().softmax(-1)).abs().max() < 2 ** -8
    lg = x.to(cuda_device).clone()
    am = ops.log_softmax_argmax_(lg)
    assert (lg.cpu() - x.log_softmax(-1)).abs().max() < 2e-6
    assert torch.equal(am.cpu().long(), x.argmax(-1))
    assert int(am[5]) == 17
    assert torch.equal(ops.argmax_rows(x.to(cuda_device)).cpu().long(), x.argmax(-1))


def test_greedy_collapse(cuda_device):
    from lcasr_b200 import ops
    blank = 9
    g = torch.Generator().manual_seed(26)
    for N in (1, 5, 1024, 1025, 5000):
        ids = torch.randint(7, 10, (3, N), generator=g).int()
        ids[1] = blank  # all blank
        tokens, n = ops.greedy_collapse(ids.to(cuda_device), blank)
        for b in range(3):
            ref = [i for i in torch.unique_consecutive(ids[b]).tolist() if i != blank]
            assert tokens[b, : int(n[b])].tolist() == ref
    ids = torch.randint(0, 4, (2, 100), generator=g).int()
    lens = torch.tensor([100, 37], dtype=torch.int32)
    tokens, n = ops.greedy_collapse(ids.to(cuda_device), 3, lens.to(cuda_device))
    ref = [i for i in torch.unique_consecutive(ids[1, :37]).tolist() if i != 3]
    assert tokens[1, : int(n[1])].tolist() == ref


@pytest.mark.parametrize("B,N,V,S", [(2, 33, 128, 9), (1, 128, 4096, 38), (3, 50, 16, 0), (2, 700, 256, 300), (1, 40, 8, 20)])
def test_ctc_loss_forward_backward(cuda_device, B, N, V, S):
    import lcasr_b200
    from lcasr_b200 import ops
    blank = V - 1
    g = torch.Generator().manual_seed(27)
    lp = torch.randn(B, N, V, generator=g).log_softmax(-1)
    Sm = max(S, 1)
    tgt = torch.randint(0, V - 1, (B, Sm), generator=g)
    if S >= 4:
        tgt[:, 2] = tgt[:, 1]  # repeated labels
    tl = torch.tensor([S if b == 0 else max(S - b, 0) for b in range(B)], dtype=torch.long)
    il = torch.tensor([N if b == 0 else N - 3 * b for b in range(B)], dtype=torch.int32)
    ref = O.ctc_loss(lp.numpy(), tgt.numpy(), il.numpy(), tl.numpy(), blank)
    nll, alpha = ops.ctc_loss_fwd(lp.to(cuda_device), tgt.to(cuda_device), il.to(cuda_device), tl.to(cuda_device), blank,
                                  keep_alpha=True)
    np.testing.assert_allclose(nll.cpu().numpy(), ref, rtol=1e-4)  # north_star: CTC loss within 1e-3 relative
    tref = F.ctc_loss(lp.transpose(0, 1), tgt, il.long(), tl, blank=blank, reduction="none")
    np.testing.assert_allclose(nll.cpu().numpy(), tref.numpy(), rtol=1e-4)
    # backward through the drop-in module, reduction='sum' as exp/train.py:104
    lpc = lp.to(cuda_device).requires_grad_(True)
    loss = lcasr_b200.CTCLoss(blank=blank, reduction="sum")(lpc.transpose(0, 1), tgt, il, tl)
    loss.backward()
    gref = O.ctc_grad(lp.numpy(), tgt.numpy(), il.numpy(), tl.numpy(), blank)
    assert abs(loss.item() - ref.sum()) < 1e-4 * abs(ref.sum())
    # fp32 log-domain recursion (like ATen's) against the float64 oracle: |alpha+beta| grows ~ N, so
    # the absolute error of exp(.) grows with the sequence length
    np.testing.assert_allclose(lpc.grad.cpu().numpy(), gref, rtol=5e-3, atol=2e-4 * max(1.0, N / 64))
    report(test="ctc", B=B, N=N, V=V, S=S, nll_rel=float(np.abs(nll.cpu().numpy() - ref).max() / np.abs(ref).max()),
           grad_abs=float(np.abs(lpc.grad.cpu().numpy() - gref).max()))


@pytest.mark.parametrize("B,N,V,S", [(2, 2500, 64, 700), (1, 7000, 128, 3000), (1, 20000, 4096, 9000)])
def test_ctc_loss_cluster_kernel_long_sequences(cuda_device, B, N, V, S):
    """> 1024 extended states: the thread-block-cluster kernel (DSMEM boundary exchange) against torch's
    CPU ctc_loss (ATen) for the loss and its gradient."""
    import lcasr_b200
    blank = V - 1
    g = torch.Generator().manual_seed(31)
    lp = torch.randn(B, N, V, generator=g).log_softmax(-1)
    tgt = torch.randint(0, V - 1, (B, S), generator=g)
    tgt[:, 5] = tgt[:, 4]
    tl = torch.tensor([S - 7 * b for b in range(B)], dtype=torch.long)
    il = torch.tensor([N - 11 * b for b in range(B)], dtype=torch.int32)
    def torch_ctc(x):
        x = x.clone().requires_grad_(True)
        loss = F.ctc_loss(x.transpose(0, 1), tgt, il.long(), tl, blank=blank, reduction="none")
        loss.sum().backward()
        return loss.detach(), x.grad
    ref64, g64 = torch_ctc(lp.double())   # ground truth
    ref32, g32 = torch_ctc(lp)            # what the reference computes (ATen, fp32 log-domain recursion)
    lpc = lp.to(cuda_device).requires_grad_(True)
    nll = lcasr_b200.CTCLoss(blank=blank, reduction="none")(lpc.transpose(0, 1), tgt, il, tl)
    nll.sum().backward()
    rel = ((nll.detach().cpu().double() - ref64).abs() / ref64.abs()).max().item()
    gerr = (lpc.grad.cpu().double() - g64).abs().max().item()
    gerr_torch32 = (g32.double() - g64).abs().max().item()
    report(test="ctc_cluster", B=B, N=N, V=V, S=S, nll_rel=rel, grad_abs_vs_fp64=gerr, torch_fp32_grad_abs_vs_fp64=gerr_torch32)
    assert rel < 1e-5  # north_star: CTC loss within 1e-3 relative
    # |alpha+beta| grows ~ 8*N, so an fp32 recursion cannot resolve the exponent better than its ulp: the
    # bar is "no worse than the reference's own fp32 implementation" (x2 margin)
    assert gerr < 2.0 * gerr_torch32 + 1e-3


@pytest.mark.parametrize("C,T,B", [(64, 264, 2), (256, 1027, 1), (512, 77, 1), (128, 8, 1)])
def test_subsampling_fused_conv0_dw(cuda_device, C, T, B):
    """conv0 + SiLU + first depthwise level in one kernel against torch fp32 and against the two
    separate kernels (same bf16 rounding point: the conv0 activation)."""
    from lcasr_b200 import ops
    Fq = 80
    spec = _rand(B, Fq, T, seed=40)
    w0, b0 = _rand(C, 1, 3, 3, seed=41), 0.1 * _rand(C, seed=42)
    w1, b1 = _rand(C, 1, 3, 3, seed=43) / 3, 0.1 * _rand(C, seed=44)
    a0 = F.silu(F.conv2d(spec.transpose(1, 2).unsqueeze(1), w0, b0, stride=2, padding=1))
    ref = F.conv2d(a0, w1, b1, stride=2, padding=1, groups=C).permute(0, 2, 3, 1)  # [B,T2,F2,C]
    args = [t.to(cuda_device) for t in (spec, w0.reshape(C, 9).contiguous(), b0, w1.reshape(C, 9).contiguous(), b1)]
    got = ops.subsample_conv0_dw(*args)
    assert got.shape == ref.shape
    assert (got.float().cpu() - ref).abs().max() < 2 ** -6 * max(1.0, ref.abs().max())
    two = ops.subsample_dwconv(ops.subsample_conv0(args[0], args[1], args[2], torch.bfloat16), args[3], args[4])
    assert (got.float() - two.float()).abs().max().item() <= 2 ** -7 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,N,V,S", [(1, 3000, 512, 1400), (3, 900, 128, 400), (8, 512, 64, 150), (1, 45, 32, 40), (2, 4000, 4096, 2100)])
def test_ctc_wavefront_is_bit_identical_to_plain_recursion(cuda_device, B, N, V, S):
    """The time-skewed wavefront form (state chunks on persistent CTAs, boundary hand-off through global memory) performs the
    same arithmetic per state as the one-CTA / cluster recursions: losses and both lattices must be bit-identical."""
    from lcasr_b200 import ops, _lib as L
    blank = V - 1
    g = torch.Generator().manual_seed(5)
    lp = torch.randn(B, N, V, generator=g).log_softmax(-1).to(cuda_device)
    tgt = torch.randint(0, V - 1, (B, S), generator=g)
    tgt[:, 3] = tgt[:, 2]
    tgt = tgt.to(cuda_device)
    tl = torch.tensor([S - 5 * b for b in range(B)], dtype=torch.long, device=cuda_device)
    il = torch.tensor([N - 7 * b for b in range(B)], dtype=torch.int32, device=cuda_device)
    assert ops.ctc_wavefront_applies(B, N, S, False)
    nll_w, alpha_w = ops.ctc_loss_fwd(lp, tgt, il, tl, blank, keep_alpha=True)
    nll_p = torch.empty_like(nll_w)
    alpha_p = torch.full_like(alpha_w, float("nan"))
    L.call("lcasr_ctc_loss_fwd", L.ptr(lp), B, N, V, L.ptr(tgt), S, L.ptr(il), L.ptr(tl), blank, L.ptr(nll_p), L.ptr(alpha_p),
           L.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(nll_w, nll_p)
    for b in range(B):  # rows / states beyond a sample's lengths are unspecified
        t, lpb = int(il[b]), 2 * int(tl[b]) + 1
        assert torch.equal(alpha_w[b, :t, :lpb], alpha_p[b, :t, :lpb])
    if 2 * S + 1 <= 4096 and ops.ctc_wavefront_applies(B, N, S, True):  # concurrent alpha / beta (training form)
        nll2, a2, b2 = ops.ctc_loss_fwd_ab(lp, tgt, il, tl, blank)
        nll3, a3, b3 = torch.empty_like(nll2), torch.empty_like(a2), torch.empty_like(b2)
        L.call("lcasr_ctc_loss_fwd_ab", L.ptr(lp), B, N, V, L.ptr(tgt), S, L.ptr(il), L.ptr(tl), blank, L.ptr(nll3), L.ptr(a3),
               L.ptr(b3), L.current_stream())
        torch.cuda.synchronize()
        assert torch.equal(nll2, nll3) and torch.equal(nll2, nll_p)
        for b in range(B):
            t, lpb = int(il[b]), 2 * int(tl[b]) + 1
            assert torch.equal(a2[b, :t, :lpb], a3[b, :t, :lpb]) and torch.equal(b2[b, :t, :lpb], b3[b, :t, :lpb])


@pytest.mark.parametrize("kind", ["layer_norm", "rms_norm"])
@pytest.mark.parametrize("M,d,n", [(300, 768, 2), (129, 256, 3), (64, 2048, 2), (7, 128, 3)])
def test_layernorm_chain_equals_sequential_norms(cuda_device, M, d, n, kind):
    """norm_out -> decoder.norm (-> decoder.norm) on rows kept in registers == the same norms applied one kernel at a time"""
    from lcasr_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = (torch.randn(M, d, generator=g) * 3 + 0.5).to(cuda_device)
    ws = [(1 + 0.2 * torch.randn(d, generator=g)).to(cuda_device) for _ in range(n)]
    bs = [(0.1 * torch.randn(d, generator=g)).to(cuda_device) if kind == "layer_norm" else None for _ in range(n)]
    eps = 1e-5 if kind == "layer_norm" else 1e-8
    cur, first = x, None
    for i in range(n):
        cur, lo = ops.layernorm(cur, ws[i], bs[i], eps, kind, out_f32=True, lo_dtype=torch.bfloat16)
        if i == 0:
            first = cur
    o32, olo = ops.layernorm_chain(x, ws, bs, eps, kind, f32_stage=0, lo_dtype=torch.bfloat16)
    # same arithmetic; the compiler may contract a*b+c differently in the two kernels: allow fp32 rounding noise
    assert (o32 - first).abs().max().item() <= 2e-6 * max(1.0, first.abs().max().item())
    diff = (olo.float() - lo.float()).abs()
    assert diff.max().item() <= 2 ** -7 * max(1.0, lo.float().abs().max().item()) and (diff > 0).float().mean().item() < 1e-3
