"""GPU: the tcgen05 / TMEM / TMA GEMM against fp32 matmul of the same bf16 operands."""
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import report

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


SHAPES = [(128, 128, 64), (128, 256, 64), (256, 256, 128), (66, 192, 64), (333, 128, 320), (1000, 4096, 256),
          (2048, 768, 3072), (4096, 2304, 768), (130, 32, 32), (20000, 256, 256), (513, 72, 40), (16384, 768, 768)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tcgen05_plain(cuda_device, M, N, K):
    from lcasr_b200 import ops, _lib as L
    a = _rand(M, K, seed=1).bfloat16()
    w = (_rand(N, K, seed=2) / math.sqrt(K)).bfloat16()
    ref = a.float().to(cuda_device) @ w.float().to(cuda_device).T
    got = ops.gemm(a.to(cuda_device), w.to(cuda_device), out_dtype=torch.float32, impl=L.GEMM_TCGEN05)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    report(test="gemm_tc", M=M, N=N, K=K, max_abs=err)
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), f"tcgen05 GEMM mismatch {err}"
    gb = ops.gemm(a.to(cuda_device), w.to(cuda_device), out_dtype=torch.bfloat16, impl=L.GEMM_TCGEN05)
    assert (gb.float() - ref).abs().max().item() < 2 ** -7 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(333, 256, 128), (2048, 768, 3072), (1000, 4096, 256)])
def test_gemm_tcgen05_epilogues(cuda_device, M, N, K):
    from lcasr_b200 import ops, _lib as L
    a = _rand(M, K, seed=3).bfloat16().to(cuda_device)
    w = (_rand(N, K, seed=4) / math.sqrt(K)).bfloat16().to(cuda_device)
    bias = _rand(N, seed=5).to(cuda_device)
    resid = _rand(M, N, seed=6).to(cuda_device)
    ref = a.float() @ w.float().T + bias
    got = ops.gemm(a, w, bias=bias, act=L.ACT_GELU_TANH, out_dtype=torch.bfloat16, impl=L.GEMM_TCGEN05)
    assert (got.float() - F.gelu(ref, approximate="tanh")).abs().max().item() < 3e-2
    got = ops.gemm(a, w, bias=bias, act=L.ACT_SILU, out_dtype=torch.float32, impl=L.GEMM_TCGEN05)
    assert (got - F.silu(ref)).abs().max().item() < 2e-3
    r = resid.clone()
    got = ops.gemm(a, w, bias=bias, resid=r, alpha=0.5, impl=L.GEMM_TCGEN05, out=r)
    assert (got - (resid + 0.5 * ref)).abs().max().item() < 2e-3
    # agreement with the independent SIMT kernel on identical operands
    simt = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, impl=L.GEMM_SIMT)
    tc = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, impl=L.GEMM_TCGEN05)
    assert (simt - tc).abs().max().item() < 1e-3


@pytest.mark.parametrize("Dh,H,N,B", [(32, 8, 300, 2), (128, 6, 257, 1), (64, 4, 1000, 1), (32, 24, 2048, 1)])
def test_fused_rotary_epilogue_and_strided_attention(cuda_device, Dh, H, N, B):
    """qkv GEMM with the rotary embedding in its epilogue (interleaved q/k head rows) + attention on the column blocks of the
    projection == plain GEMM -> rope_split -> attention (the unfused kernels), and both match an fp32 torch statement."""
    import torch.nn.functional as F
    from lcasr_b200 import ops, _lib as L
    d = H * Dh
    g = torch.Generator().manual_seed(9)
    a = torch.randn(B * N, d, generator=g).bfloat16().to(cuda_device)
    w = (torch.randn(3 * d, d, generator=g) / d ** 0.5).bfloat16()
    qk = w[: 2 * d].reshape(2 * H, Dh, d)
    w_il = torch.cat([torch.stack([qk[:, : Dh // 2], qk[:, Dh // 2:]], 2).reshape(2 * d, d), w[2 * d:]], 0).contiguous().to(cuda_device)
    w = w.to(cuda_device)
    inv_freq = (1.0 / (1500000 ** (torch.arange(0, Dh, 2).float() / Dh))).to(cuda_device)
    cos, sin = ops.rope_table(inv_freq, 1.0, N)
    cos_t, sin_t = ops.rope_table(inv_freq, 1.0, N, transposed=True)
    assert torch.equal(cos_t.t().contiguous(), cos) and torch.equal(sin_t.t().contiguous(), sin)
    fused = ops.attention_qkv(ops.gemm_rope(a, w_il, cos_t, sin_t, N, 2 * d, Dh), B, N, H, Dh)
    q, k, v = ops.rope_split(ops.gemm(a, w), B, N, H, Dh, cos, sin)
    unfused = ops.attention(q, k, v)
    err = (fused.float() - unfused.float()).abs().max().item()
    # fp32 statement: rotate_half form of rotary_emb.py:61-73 on the fp32 projection
    qkv32 = (a.float() @ w.float().t()).view(B, N, 3, H, Dh)
    c2, s2 = torch.cat([cos, cos], -1)[None, :, None, :], torch.cat([sin, sin], -1)[None, :, None, :]
    rot = lambda x: torch.cat([-x[..., Dh // 2:], x[..., : Dh // 2]], -1)
    q32, k32, v32 = qkv32[:, :, 0], qkv32[:, :, 1], qkv32[:, :, 2]
    q32, k32 = q32 * c2 + rot(q32) * s2, k32 * c2 + rot(k32) * s2
    ref = F.scaled_dot_product_attention(q32.transpose(1, 2), k32.transpose(1, 2), v32.transpose(1, 2)).transpose(1, 2).reshape(B, N, d)
    e_f, e_u = (fused.float() - ref).abs().max().item(), (unfused.float() - ref).abs().max().item()
    report(test="fused_rope_qkv", Dh=Dh, H=H, N=N, fused_vs_unfused=err, fused_vs_fp32=e_f, unfused_vs_fp32=e_u)
    assert e_f < 2e-2 and err < 2e-2
    assert e_f < 1.5 * e_u + 2e-3  # rotating in fp32 BEFORE the bf16 rounding is at least as accurate as rounding first


@pytest.mark.parametrize("M,d", [(700, 256), (4096, 768), (130, 2048), (33, 384)])
def test_fused_glu_epilogue(cuda_device, M, d):
    """pointwise_conv1 + GLU in one kernel (packed value/gate rows) == GEMM -> glu kernel, and the fp32 statement."""
    from lcasr_b200 import ops
    g = torch.Generator().manual_seed(10)
    a = torch.randn(M, d, generator=g).bfloat16().to(cuda_device)
    w = (torch.randn(2 * d, d, generator=g) / d ** 0.5).bfloat16()
    b = 0.3 * torch.randn(2 * d, generator=g)
    w_glu = torch.stack([w[:d].reshape(d // 32, 32, d), w[d:].reshape(d // 32, 32, d)], 1).reshape(2 * d, d).contiguous().to(cuda_device)
    b_glu = torch.stack([b[:d].reshape(d // 32, 32), b[d:].reshape(d // 32, 32)], 1).reshape(2 * d).contiguous().to(cuda_device)
    fused = ops.gemm_glu(a, w_glu, b_glu)
    unfused = ops.glu(ops.gemm(a, w.to(cuda_device), bias=b.to(cuda_device)))
    y = a.float().cpu() @ w.float().t() + b
    ref = y[:, :d] * torch.sigmoid(y[:, d:])
    e_f, e_u = (fused.float().cpu() - ref).abs().max().item(), (unfused.float().cpu() - ref).abs().max().item()
    report(test="fused_glu", M=M, d=d, fused_vs_fp32=e_f, unfused_vs_fp32=e_u)
    assert fused.shape == (M, d)
    assert e_f < 2 ** -7 * max(1.0, ref.abs().max().item()) and e_f < 1.5 * e_u + 1e-3
