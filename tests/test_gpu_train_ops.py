"""GPU: every kernel of the training step (csrc/{gemm_tcx,train,subsample_bwd}.cu) against torch fp32 / autograd
on the same bf16-rounded inputs.  Calls go through the C ABI (ctypes)."""
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import report

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _close(got, ref, rel, what):
    err = (got.float() - ref.float()).abs().max().item()
    den = max(1e-6, ref.float().abs().max().item())
    report(test="train_op", what=what, rel_err=err / den)
    assert err <= rel * den, f"{what}: max-abs {err} vs scale {den}"


# ---- gemm_ex: operand majorness, split-K accumulation, batching, epilogues ------------------------------------

@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (333, 72, 200), (1000, 768, 3072), (2048, 3072, 768), (4096, 256, 264)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_ex_majorness(cuda_device, M, N, K, a_mn, b_mn):
    from lcasr_b200 import train_ops as T
    if a_mn and M % 8:
        M = (M + 7) // 8 * 8  # a stored [K, M] matrix needs a 16-byte row pitch
    a = _rand(M, K, seed=1).to(BF).to(cuda_device)
    b = _rand(N, K, seed=2, scale=1 / math.sqrt(K)).to(BF).to(cuda_device)
    ref = a.float() @ b.float().T
    A = a.t().contiguous() if a_mn else a
    Bm = b.t().contiguous() if b_mn else b
    Kp = (K + 7) // 8 * 8
    if not a_mn and K % 8:
        A = F.pad(a, (0, Kp - K)).contiguous()
    if not b_mn and K % 8:
        Bm = F.pad(b, (0, Kp - K)).contiguous()
    out = torch.empty(M, N, dtype=BF, device=cuda_device)
    T.gemm_ex(A, Bm, out, M, N, K, a_mn=a_mn, b_mn=b_mn, lda=A.shape[1], ldb=Bm.shape[1], ldo=N)
    _close(out, ref, 2 ** -7, f"gemm_ex bf16 a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}")
    acc = torch.ones(M, N, dtype=torch.float32, device=cuda_device)
    T.gemm_ex(A, Bm, acc, M, N, K, a_mn=a_mn, b_mn=b_mn, lda=A.shape[1], ldb=Bm.shape[1], ldo=N, alpha=0.5)
    _close(acc, 1.0 + 0.5 * ref, 1e-3, f"gemm_ex f32 accumulate a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}")


@pytest.mark.parametrize("M,No,Ni", [(16384, 768, 768), (4096, 3072, 768), (2048, 768, 4096), (520, 64, 320)])
def test_dgrad_wgrad(cuda_device, M, No, Ni):
    from lcasr_b200 import train_ops as T
    dy = _rand(M, No, seed=3).to(BF).to(cuda_device)
    x = _rand(M, Ni, seed=4).to(BF).to(cuda_device)
    w = _rand(No, Ni, seed=5, scale=1 / math.sqrt(No)).to(BF).to(cuda_device)
    _close(T.dgrad(dy, w), dy.float() @ w.float(), 2 ** -7, f"dgrad {M}x{No}x{Ni}")
    dw = torch.zeros(No, Ni, dtype=torch.float32, device=cuda_device)
    T.wgrad(dy, x, dw)
    T.wgrad(dy, x, dw, alpha=0.5)  # accumulates
    _close(dw, 1.5 * (dy.float().T @ x.float()), 2e-3, f"wgrad split-K {M}x{No}x{Ni}")


@pytest.mark.parametrize("epi", ["gelu", "silu"])
def test_gemm_ex_activation_backward_epilogues(cuda_device, epi):
    from lcasr_b200 import train_ops as T, _lib as L
    M, No, Ni = 1000, 256, 1024
    dy = _rand(M, No, seed=6).to(BF).to(cuda_device)
    w = _rand(No, Ni, seed=7, scale=1 / math.sqrt(No)).to(BF).to(cuda_device)
    pre = _rand(M, Ni, seed=8, scale=2.0).to(BF).to(cuda_device)
    p32 = pre.float().requires_grad_(True)
    y = F.gelu(p32, approximate="tanh") if epi == "gelu" else F.silu(p32)
    y.backward(0.5 * (dy.float() @ w.float()))
    got = T.dgrad(dy, w, aux=pre, epi=L.EPI_GELU_BWD if epi == "gelu" else L.EPI_SILU_BWD, alpha=0.5)
    _close(got, p32.grad, 2 ** -6, f"dgrad+{epi}'")


@pytest.mark.parametrize("B,N,H,Dh", [(2, 256, 2, 128), (3, 200, 4, 32), (1, 1000, 2, 64), (8, 2048, 6, 128), (2, 33, 2, 32),
                                        (2, 65, 1, 128), (1, 64, 1, 64), (2, 129, 3, 64), (1, 4100, 2, 128)])
def test_attention_train_and_backward(cuda_device, B, N, H, Dh):
    from lcasr_b200 import train_ops as T
    q = _rand(B, N, H, Dh, seed=11).to(BF).to(cuda_device)
    k = _rand(B, N, H, Dh, seed=12).to(BF).to(cuda_device)
    v = _rand(B, N, H, Dh, seed=13).to(BF).to(cuda_device)
    do = _rand(B, N, H * Dh, seed=14).to(BF).to(cuda_device)
    out, lse2 = T.attention_train(q, k, v)
    qf, kf, vf = (t.float().transpose(1, 2).requires_grad_(True) for t in (q, k, v))  # [B,H,N,Dh]
    s = (qf @ kf.transpose(-1, -2)) / math.sqrt(Dh)
    ref_lse2 = torch.logsumexp(s, dim=-1) * 1.4426950408889634
    ref = F.scaled_dot_product_attention(qf, kf, vf)
    _close(out.view(B, N, H, Dh), ref.transpose(1, 2), 2 ** -6, f"attention_train out {B},{N},{H},{Dh}")
    assert (lse2 - ref_lse2).abs().max().item() < 2e-2, "log-sum-exp (bf16 P row sums)"
    ref.backward(do.float().view(B, N, H, Dh).transpose(1, 2))
    dq, dk, dv = T.attention_bwd(q, k, v, out.view(B, N, H, Dh), do.view(B, N, H, Dh), lse2)
    for name, got, r in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        _close(got, r.transpose(1, 2), 3e-2, f"attention_bwd {name} {B},{N},{H},{Dh}")
    # the materialised forms (fused P/dS kernel + 3 batched GEMMs; two-GEMM formulation) against each other and, for the head
    # dims the flash-style kernels cover, against those (dq, dk, dv above are the flash result there)
    dq1, dk1, dv1 = T.attention_bwd(q, k, v, out.view(B, N, H, Dh), do.view(B, N, H, Dh), lse2, flash=False)
    dq3, dk3, dv3 = T.attention_bwd(q, k, v, out.view(B, N, H, Dh), do.view(B, N, H, Dh), lse2, fused_pds=False, flash=False)
    for a_, b_, c_ in ((dq, dq3, dq1), (dk, dk3, dk1), (dv, dv3, dv1)):
        assert (c_.float() - b_.float()).abs().max().item() <= 2 ** -6 * b_.float().abs().max().item()
        assert (a_.float() - b_.float()).abs().max().item() <= 2 ** -5 * b_.float().abs().max().item()
    if T.flash_applies(Dh):  # repeatable bit for bit (no atomics anywhere)
        dq4, dk4, dv4 = T.attention_bwd(q, k, v, out.view(B, N, H, Dh), do.view(B, N, H, Dh), lse2)
        assert torch.equal(dq, dq4) and torch.equal(dk, dk4) and torch.equal(dv, dv4)
    if B > 1:  # chunking over recordings gives the same result
        dq2, dk2, dv2 = T.attention_bwd(q, k, v, out.view(B, N, H, Dh), do.view(B, N, H, Dh), lse2, chunk_b=1, flash=False)
        assert torch.equal(dq1, dq2) and torch.equal(dk1, dk2) and torch.equal(dv1, dv2)


@pytest.mark.parametrize("B,N,H,Dh,lens", [(3, 200, 2, 32, [200, 77, 136]), (2, 300, 1, 128, [129, 300]), (4, 2048, 6, 128, [2048, 1000, 2047, 5])])
def test_attention_train_padded_batch(cuda_device, B, N, H, Dh, lens):
    """key-padding mask + zeroed rows of padded tokens (attention.py:511,541 in a padded batch) against SDPA with the same
    masks; the backward treats every recording as the n_b x n_b problem of its valid tokens"""
    from lcasr_b200 import train_ops as T
    lens_dev = torch.tensor(lens, dtype=torch.int32, device=cuda_device)
    pad = (torch.arange(N)[None, :] >= torch.tensor(lens)[:, None]).to(cuda_device)        # [B,N]
    q, k, v = (_rand(B, N, H, Dh, seed=s_).to(BF).to(cuda_device) for s_ in (21, 22, 23))
    for t in (q, k, v):
        T.mask_rows_(t, lens_dev, B, N)                                                    # qkv of zeroed rows (no bias) is zero
        assert float(t[pad].abs().max()) == 0 and float(t[~pad].abs().min()) >= 0
    do = _rand(B, N, H * Dh, seed=24).to(BF).to(cuda_device)
    out, lse2 = T.attention_train(q, k, v, lens_dev)
    T.mask_rows_(out, lens_dev, B, N)
    qf, kf, vf = (t.float().transpose(1, 2).requires_grad_(True) for t in (q, k, v))
    bias = torch.zeros(B, 1, 1, N, device=cuda_device).masked_fill(pad[:, None, None, :], float("-inf"))
    ref = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=bias).transpose(1, 2)      # [B,N,H,Dh]
    ref = ref.masked_fill(pad[:, :, None, None], 0.0)
    _close(out.view(B, N, H, Dh), ref, 2 ** -6, "masked attention_train out")
    ref.backward(do.float().view(B, N, H, Dh))
    for fused, flash in ((True, None), (True, False), (False, False)):
        dq, dk, dv = T.attention_bwd(q, k, v, out.view(B, N, H, Dh), do.view(B, N, H, Dh), lse2, lens=lens, fused_pds=fused, flash=flash)
        for name, got, r in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
            _close(got, r.transpose(1, 2), 3e-2, f"padded attention_bwd {name} fused={fused} flash={flash}")
            assert float(got[pad].abs().max()) == 0, f"{name}: rows of padded tokens must get no gradient"


# ---- memory-bound backward kernels ----------------------------------------------------------------------------

def test_elementwise_backward_kernels(cuda_device):
    from lcasr_b200 import train_ops as T, _lib as L
    dev = cuda_device
    M, d = 777, 256
    x = _rand(M, d, seed=20).to(dev)
    _close(T.scale_cast(x, 0.5), 0.5 * x, 2 ** -8, "scale_cast")
    xb = x.to(BF)
    _close(T.act_fwd(xb, L.ACT_GELU_TANH), F.gelu(xb.float(), approximate="tanh"), 2 ** -7, "act_fwd gelu")
    _close(T.act_fwd(xb, L.ACT_SILU), F.silu(xb.float()), 2 ** -7, "act_fwd silu")
    acc = x.clone()
    _close(T.add_bf16_(acc, xb), x + xb.float(), 1e-6, "add_bf16")
    # GLU
    u = _rand(M, 2 * d, seed=21).to(BF).to(dev)
    dg = _rand(M, d, seed=22).to(BF).to(dev)
    u32 = u.float().requires_grad_(True)
    F.glu(u32, dim=-1).backward(dg.float())
    _close(T.glu_bwd(u, dg), u32.grad, 2 ** -6, "glu_bwd")
    # softmax / log-softmax
    V = 4096
    logits = _rand(300, V, seed=23, scale=3.0).to(dev)
    p = logits.softmax(-1).to(BF)
    dp = _rand(300, V, seed=24).to(BF).to(dev)
    ref = p.float() * (dp.float() - (p.float() * dp.float()).sum(-1, keepdim=True))
    got = T.softmax_bwd(p, dp)
    assert (got.float() - ref).abs().max().item() < 2 ** -7 * ref.abs().max().item() + 1e-6
    lp = logits.log_softmax(-1)
    dlp = _rand(300, V, seed=25).to(dev)
    l32 = logits.clone().requires_grad_(True)
    l32.log_softmax(-1).backward(dlp)
    _close(T.log_softmax_bwd(lp, dlp, 1.0), l32.grad, 2 ** -7, "log_softmax_bwd")
    # column sums
    out = torch.ones(d, device=dev)
    T.colsum_(out, xb, 2.0)
    _close(out, 1.0 + 2.0 * xb.float().sum(0), 1e-4, "colsum bf16")
    out = torch.zeros(d, device=dev)
    T.colsum_(out, x)
    _close(out, x.sum(0), 1e-4, "colsum f32")


@pytest.mark.parametrize("H,Dh,rot", [(2, 32, True), (6, 128, True), (3, 64, False)])
def test_rope_backward_and_rowdot(cuda_device, H, Dh, rot):
    from lcasr_b200 import train_ops as T, ops
    dev = cuda_device
    B, N = 2, 130
    d = H * Dh
    inv_freq = (1.0 / (1.5e6 ** (torch.arange(0, Dh, 2).float() / Dh))).to(dev)
    cos, sin = ops.rope_table(inv_freq, 1.0, N) if rot else (None, None)
    qkv = _rand(B * N, 3 * d, seed=30).to(BF).to(dev)
    # reference: rope_split forward as a linear map, its transpose via autograd on the same math
    x = qkv.float().requires_grad_(True)
    q, k, v = x.view(B, N, 3, H, Dh).unbind(2)
    if rot:
        c = torch.cat([cos, cos], -1)[None, :, None, :]
        s = torch.cat([sin, sin], -1)[None, :, None, :]
        rh = lambda t: torch.cat((-t[..., Dh // 2:], t[..., : Dh // 2]), -1)
        q, k = q * c + rh(q) * s, k * c + rh(k) * s
    gq, gk, gv = (_rand(B, N, H, Dh, seed=31 + i).to(BF).to(dev) for i in range(3))
    (q * gq.float() + k * gk.float() + v * gv.float()).sum().backward()
    _close(T.rope_bwd_merge(gq, gk, gv, cos, sin), x.grad, 2 ** -7, f"rope_bwd_merge H={H} Dh={Dh} rot={rot}")
    D = torch.empty(B, H, N, device=dev)
    from lcasr_b200 import _lib as L
    L.call("lcasr_rowdot", gq.data_ptr(), gk.data_ptr(), B, N, H, Dh, D.data_ptr(), L.current_stream())
    _close(D, (gq.float() * gk.float()).sum(-1).permute(0, 2, 1), 1e-5, "rowdot")


@pytest.mark.parametrize("kind,d,dy_dtype", [("layer_norm", 768, BF), ("layer_norm", 64, torch.float32), ("rms_norm", 256, BF),
                                             ("layer_norm", 2048, BF)])
def test_layernorm_backward(cuda_device, kind, d, dy_dtype):
    from lcasr_b200 import train_ops as T
    dev = cuda_device
    M = 1037
    x = (_rand(M, d, seed=40) * 2 + 0.3).to(dev)
    w = (1 + 0.2 * _rand(d, seed=41)).to(dev)
    b = (0.1 * _rand(d, seed=42)).to(dev)
    dy = _rand(M, d, seed=43).to(dy_dtype).to(dev)
    x32, w32, b32 = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    if kind == "layer_norm":
        y = F.layer_norm(x32, (d,), w32, b32, 1e-5)
    else:
        y = w32 * (x32 / (x32.norm(2, dim=-1, keepdim=True) * d ** -0.5 + 1e-8))
    y.backward(dy.float())
    dx = torch.ones(M, d, device=dev)
    dw, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    T.layernorm_bwd(x, dy, w, dx, dw, db if kind == "layer_norm" else None, eps=1e-5 if kind == "layer_norm" else 1e-8,
                    kind=kind, accumulate=True)
    _close(dx, 1.0 + x32.grad, 1e-4, f"{kind} dx d={d}")
    _close(dw, w32.grad, 1e-3, f"{kind} dweight d={d}")
    if kind == "layer_norm":
        _close(db, b32.grad, 1e-3, f"{kind} dbias d={d}")
    T.layernorm_bwd(x, dy, w, dx, dw, db if kind == "layer_norm" else None, eps=1e-5 if kind == "layer_norm" else 1e-8,
                    kind=kind, accumulate=False)
    _close(dx, x32.grad, 1e-4, f"{kind} dx (overwrite) d={d}")


@pytest.mark.parametrize("B,N,d,ks", [(2, 300, 64, 9), (8, 2048, 768, 9), (1, 77, 256, 15)])
def test_conv_module_training_kernels(cuda_device, B, N, d, ks):
    """depthwise conv + BatchRenorm(train) + SiLU forward and backward against autograd through the reference's
    formulas (batchrenorm.py:52-84, convolution.py:112-121)."""
    from lcasr_b200 import train_ops as T
    dev = cuda_device
    g = _rand(B, N, d, seed=50).to(BF).to(dev)
    w = _rand(d, ks, seed=51, scale=0.3).to(dev)
    b = _rand(d, seed=52, scale=0.1).to(dev)
    bw = (1 + 0.2 * _rand(d, seed=53)).to(dev)
    bb = (0.1 * _rand(d, seed=54)).to(dev)
    rm = (0.05 * _rand(d, seed=55)).to(dev)
    rs = (1 + 0.1 * _rand(d, seed=56).abs()).to(dev)
    eps, rmax, dmax, mom = 1e-3, 1.5, 0.3, 0.01
    dy = _rand(B, N, d, seed=57).to(BF).to(dev)

    c, sums = T.dwconv1d_fwd(g, w, b, stats=True)
    g32, w32, b32 = g.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    c_ref = F.conv1d(g32.transpose(1, 2), w32[:, None, :], b32, padding=(ks - 1) // 2, groups=d).transpose(1, 2)
    _close(c, c_ref, 2 ** -7, "dwconv1d_fwd")
    # BatchRenorm train on OUR rounded c (so the two sides see the same input)
    c32 = c.float().requires_grad_(True)
    bw32, bb32 = bw.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    flat = c32.reshape(-1, d)
    mu, sd = flat.mean(0), flat.std(0, unbiased=False) + eps
    r = (sd.detach() / rs).clamp(1 / rmax, rmax)
    dd = ((mu.detach() - rm) / rs).clamp(-dmax, dmax)
    z = bw32 * ((c32 - mu) / sd * r + dd) + bb32
    y_ref = F.silu(z)
    rm2, rs2 = rm.clone(), rs.clone()
    A, Bc, stats = T.brn_train_stats(sums, B * N, rm2, rs2, eps, rmax, dmax, mom, bw, bb)
    _close(stats[0], mu, 1e-3, "brn mean")
    _close(stats[1], sd, 1e-3, "brn sigma")
    _close(rm2, rm + mom * (mu.detach() - rm), 1e-4, "running_mean update")
    _close(rs2, rs + mom * (sd.detach() - rs), 1e-4, "running_std update")
    y = T.affine_silu(c, A, Bc)
    _close(y, y_ref, 2 ** -6, "brn+silu forward")
    y_ref.backward(dy.float())
    dbw, dbb = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    dc = T.brn_silu_bwd(c, dy, A, Bc, stats, bw, dbw, dbb)
    _close(dc, c32.grad, 3e-2, "brn+silu dc")
    _close(dbw, bw32.grad, 1e-2, "brn dweight")
    _close(dbb, bb32.grad, 1e-2, "brn dbias")
    # depthwise conv backward for an arbitrary upstream gradient
    c_ref.backward(dy.float())
    _close(T.dwconv1d_bwd_data(dy, w), g32.grad, 2 ** -7, "dwconv1d_bwd_data")
    dw, db = torch.zeros(d, ks, device=dev), torch.zeros(d, device=dev)
    T.dwconv1d_bwd_weight_(g, dy, dw, db)
    _close(dw, w32.grad, 2e-3, "dwconv1d dweight")
    _close(db, b32.grad, 2e-3, "dwconv1d dbias")


@pytest.mark.parametrize("B,T_,C", [(2, 100, 32), (1, 333, 64), (2, 1024, 256)])
def test_subsampling_backward_kernels(cuda_device, B, T_, C):
    from lcasr_b200 import train_ops as T, ops
    dev = cuda_device
    Fdim = 80
    spec = _rand(B, Fdim, T_, seed=60).to(dev)
    w0 = _rand(C, 9, seed=61, scale=0.3).to(dev)
    b0 = _rand(C, seed=62, scale=0.1).to(dev)
    w1 = _rand(C, 9, seed=63, scale=0.3).to(dev)
    b1 = _rand(C, seed=64, scale=0.1).to(dev)
    s1 = ops.subsample_conv0(spec, w0, b0, out_dtype=BF)            # [B,T1,40,C]
    s2 = ops.subsample_dwconv(s1, w1, b1)                            # [B,T2,20,C]
    ds2 = _rand(*s2.shape, seed=65).to(BF).to(dev)
    # autograd reference in the reference's NCHW layout
    w0r, b0r = w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    w1r, b1r = w1.clone().requires_grad_(True), b1.clone().requires_grad_(True)
    img = spec.transpose(1, 2).unsqueeze(1)
    a1 = F.silu(F.conv2d(img, w0r.view(C, 1, 3, 3), b0r, stride=2, padding=1))
    a1.retain_grad()
    s1_32 = s1.float().permute(0, 3, 1, 2).requires_grad_(True)      # the rounded activation our kernels read
    a2 = F.conv2d(s1_32, w1r.view(C, 1, 3, 3), b1r, stride=2, padding=1, groups=C)
    a2.backward(ds2.float().permute(0, 3, 1, 2))
    ds1 = T.subsample_dwconv_bwd_data(ds2, w1, s1.shape[1], s1.shape[2])
    _close(ds1, s1_32.grad.permute(0, 2, 3, 1), 2 ** -7, "subsample dwconv bwd data")
    dw1, db1 = torch.zeros(C, 9, device=dev), torch.zeros(C, device=dev)
    T.subsample_dwconv_bwd_weight_(s1, ds2, dw1, db1)
    _close(dw1, w1r.grad, 2e-3, "subsample dwconv dweight")
    _close(db1, b1r.grad, 2e-3, "subsample dwconv dbias")
    a1.backward(ds1.float().permute(0, 3, 1, 2))
    dw0, db0 = torch.zeros(C, 9, device=dev), torch.zeros(C, device=dev)
    T.subsample_conv0_bwd_(spec, w0, b0, ds1, dw0, db0)
    _close(dw0, w0r.grad, 5e-3, "conv0 dweight")
    _close(db0, b0r.grad, 5e-3, "conv0 dbias")


@pytest.mark.parametrize("B,T_,C,Fdim", [(2, 100, 64, 80), (1, 333, 64, 80), (2, 1024, 256, 80), (1, 130, 128, 37), (3, 61, 64, 18)])
def test_subsampling_level1_fused_backward(cuda_device, B, T_, C, Fdim):
    """conv0 + SiLU + depthwise level 1: the fused forward / fused backward pair (the conv0 activation never reaches HBM)
    against autograd over the same chain in fp32, and against the unfused kernels"""
    from lcasr_b200 import train_ops as T, ops
    dev = cuda_device
    spec = _rand(B, Fdim, T_, seed=70).to(dev)
    w0 = _rand(C, 9, seed=71, scale=0.3).to(dev)
    b0 = _rand(C, seed=72, scale=0.1).to(dev)
    w1 = _rand(C, 9, seed=73, scale=0.3).to(dev)
    b1 = _rand(C, seed=74, scale=0.1).to(dev)
    d1 = ops.subsample_conv0_dw(spec, w0, b0, w1, b1)               # fused forward [B,T2,F2,C]
    dd1 = _rand(*d1.shape, seed=75).to(BF).to(dev)
    w0r, b0r, w1r, b1r = (t.clone().requires_grad_(True) for t in (w0, b0, w1, b1))
    img = spec.transpose(1, 2).unsqueeze(1)
    a1 = F.silu(F.conv2d(img, w0r.view(C, 1, 3, 3), b0r, stride=2, padding=1))
    a2 = F.conv2d(a1, w1r.view(C, 1, 3, 3), b1r, stride=2, padding=1, groups=C)
    _close(d1, a2.permute(0, 2, 3, 1), 2 ** -6, "fused level-1 forward")
    a2.backward(dd1.float().permute(0, 3, 1, 2))
    dw0, db0, dw1, db1 = (torch.zeros(C, 9, device=dev), torch.zeros(C, device=dev), torch.zeros(C, 9, device=dev),
                          torch.zeros(C, device=dev))
    T.subsample_l1_bwd_(spec, w0, b0, w1, dd1, dw0, db0, dw1, db1)
    _close(dw1, w1r.grad, 4e-3, "fused l1: depthwise dweight")
    _close(db1, b1r.grad, 2e-3, "fused l1: depthwise dbias")
    _close(dw0, w0r.grad, 6e-3, "fused l1: conv0 dweight")
    _close(db0, b0r.grad, 6e-3, "fused l1: conv0 dbias")
    # accumulation semantics (+=) and agreement with the unfused kernel chain
    T.subsample_l1_bwd_(spec, w0, b0, w1, dd1, dw0, db0, dw1, db1)
    _close(dw0, 2 * w0r.grad, 6e-3, "fused l1: accumulates")
    s1 = ops.subsample_conv0(spec, w0, b0, out_dtype=BF)
    uw1, ub1, uw0, ub0 = (torch.zeros(C, 9, device=dev), torch.zeros(C, device=dev), torch.zeros(C, 9, device=dev),
                          torch.zeros(C, device=dev))
    T.subsample_dwconv_bwd_weight_(s1, dd1, uw1, ub1)
    ds1 = T.subsample_dwconv_bwd_data(dd1, w1, s1.shape[1], s1.shape[2])
    T.subsample_conv0_bwd_(spec, w0, b0, ds1, uw0, ub0)
    _close(dw1, 2 * uw1, 2e-3, "fused vs unfused: depthwise dweight")
    _close(dw0, 2 * uw0, 6e-3, "fused vs unfused: conv0 dweight (unfused rounds ds1 to bf16)")
