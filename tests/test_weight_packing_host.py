"""CPU: the two weight packings behind the fused GEMM epilogues are re-arrangements under which the kernels' arithmetic
(restated here in plain torch) equals the reference's: rotary on interleaved pairs (rotary_emb.py:44-73) and F.glu
(convolution.py:107-108).  The GPU tests compare the kernels with the same statements (tests/test_gpu_gemm_tc.py)."""
import torch

from lcasr_b200.model import pack_glu_blocks, pack_qkv_rotary_interleaved


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def test_interleaved_rotary_gives_the_reference_attention_scores():
    torch.manual_seed(0)
    N, H, Dh, d = 37, 3, 16, 48
    x = torch.randn(N, d, dtype=torch.float64)
    w = torch.randn(3 * H * Dh, d, dtype=torch.float64)        # rows already in [q | k | v] order
    inv_freq = 1.0 / (1.5e6 ** (torch.arange(0, Dh, 2, dtype=torch.float64) / Dh))
    ang = torch.arange(N, dtype=torch.float64)[:, None] * inv_freq[None, :]     # [N, Dh/2]
    cos, sin = torch.cat([ang, ang], -1).cos(), torch.cat([ang, ang], -1).sin()  # rotary_emb.py:52-55
    # reference: rotate_half on the natural head layout
    qkv = (x @ w.T).reshape(N, 3, H, Dh)
    q_ref = qkv[:, 0] * cos[:, None] + _rotate_half(qkv[:, 0]) * sin[:, None]
    k_ref = qkv[:, 1] * cos[:, None] + _rotate_half(qkv[:, 1]) * sin[:, None]
    s_ref = torch.einsum("nhd,mhd->hnm", q_ref, k_ref)
    # kernel statement: interleaved rows -> adjacent columns (2i, 2i+1) form the pair with angle i
    y = (x @ pack_qkv_rotary_interleaved(w, H, Dh).T).reshape(N, 3, H, Dh)
    c, s_ = ang.cos()[:, None, :], ang.sin()[:, None, :]       # [N, 1, Dh/2]: one (cos, sin) per pair

    def rot(t):
        a, b = t[..., 0::2], t[..., 1::2]
        out = torch.empty_like(t)
        out[..., 0::2] = a * c - b * s_
        out[..., 1::2] = b * c + a * s_
        return out
    q_il, k_il = rot(y[:, 0]), rot(y[:, 1])
    assert torch.allclose(torch.einsum("nhd,mhd->hnm", q_il, k_il), s_ref, atol=1e-9)
    assert torch.equal(y[:, 2], qkv[:, 2])                     # v untouched
    # and the interleaved q IS the reference q under the pair permutation (new 2i <- old i, new 2i+1 <- old i + Dh/2)
    perm = torch.stack([torch.arange(Dh // 2), torch.arange(Dh // 2) + Dh // 2], 1).reshape(-1)
    assert torch.allclose(q_il, q_ref[..., perm], atol=1e-9)


def test_glu_blocks_give_f_glu():
    torch.manual_seed(1)
    M, d = 19, 96
    x = torch.randn(M, d, dtype=torch.float64)
    w1, b1 = torch.randn(2 * d, d, dtype=torch.float64), torch.randn(2 * d, dtype=torch.float64)
    ref = torch.nn.functional.glu(x @ w1.T + b1, dim=-1)       # value = first d channels, gate = last d (convolution.py:108)
    wg, bg = pack_glu_blocks(w1, b1)
    y = (x @ wg.T + bg).reshape(M, d // 32, 2, 32)             # 64-column blocks: 32 values, their 32 gates
    got = (y[:, :, 0] * torch.sigmoid(y[:, :, 1])).reshape(M, d)
    assert torch.allclose(got, ref, atol=1e-12)
