"""CPU statements (fp64 torch) of the two attention algorithms the round-2 kernels implement, checked against autograd /
plain softmax attention.  They pin the ALGORITHMS — the chunking, the transposed score tiles, the packed per-row statistics
with their padding, the exact merge of key-range partial results; the kernels themselves are compared with torch on the GPU
(tests/test_gpu_train_ops.py, tests/test_gpu_scale.py, tests/test_gpu_seqpar.py).

  attn_bwd_flash_kernel (csrc/gemm_tcx.cu):   dK/dV pass over 128-key tiles x 64-query chunks with S^T / dP^T,
                                              dQ pass over 128-query tiles x 64-key chunks; statistics {lse2, D*scale}
  attn_merge_kernel (csrc/seqpar.cu) + the last-wave split of model.cu:  out = sum_s 2^(lse_s - L) O_s
"""
import math

import pytest
import torch

LOG2E = 1.4426950408889634


def _flash_backward_statement(q, k, v, o, do, lse2):
    """q, k, v, o, do: [N, Dh] (one head); lse2: [N] log2-domain log-sum-exp of the scaled scores."""
    N, Dh = q.shape
    scale = 1.0 / math.sqrt(Dh)
    c = scale * LOG2E
    Npad = (N + 127) // 128 * 128
    # attn_bwd_prep_kernel: {lse2, D * scale}, padded with {+inf, 0} so that tokens that do not exist get P = dS = 0
    stats = torch.zeros(Npad, 2, dtype=q.dtype)
    stats[:, 0] = float("inf")
    stats[:N, 0] = lse2
    stats[:N, 1] = (do * o).sum(-1) * scale

    def rows(t, r0, n):  # TMA box with zero fill beyond the tensor
        out = torch.zeros(n, Dh, dtype=t.dtype)
        m = max(0, min(n, N - r0))
        out[:m] = t[r0:r0 + m]
        return out

    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    n_chunks = (N + 63) // 64
    for j0 in range(0, N, 128):          # MODE 0: a CTA owns 128 keys, lanes = keys
        Kj, Vj = rows(k, j0, 128), rows(v, j0, 128)
        acc_v, acc_k = torch.zeros(128, Dh, dtype=q.dtype), torch.zeros(128, Dh, dtype=q.dtype)
        for ci in range(n_chunks):
            Qi, dOi = rows(q, ci * 64, 64), rows(do, ci * 64, 64)
            st = stats[ci * 64:ci * 64 + 64]
            St, dPt = Kj @ Qi.T, Vj @ dOi.T                      # [128 keys, 64 queries]
            Pt = torch.exp2(St * c - st[None, :, 0])
            dSt = Pt * (dPt * scale - st[None, :, 1])
            acc_v += Pt @ dOi
            acc_k += dSt @ Qi
        m = min(128, N - j0)
        dv[j0:j0 + m], dk[j0:j0 + m] = acc_v[:m], acc_k[:m]
    for i0 in range(0, N, 128):          # MODE 1: a CTA owns 128 queries, lanes = queries
        Qi, dOi = rows(q, i0, 128), rows(do, i0, 128)
        st = stats[i0:i0 + 128]
        acc_q = torch.zeros(128, Dh, dtype=q.dtype)
        for cj in range(n_chunks):
            Kj, Vj = rows(k, cj * 64, 64), rows(v, cj * 64, 64)
            S, dP = Qi @ Kj.T, dOi @ Vj.T
            dS = torch.exp2(S * c - st[:, None, 0]) * (dP * scale - st[:, None, 1])
            acc_q += dS @ Kj                                       # zero-filled key rows contribute nothing
        m = min(128, N - i0)
        dq[i0:i0 + m] = acc_q[:m]
    return dq, dk, dv


@pytest.mark.parametrize("N,Dh", [(64, 16), (129, 8), (200, 32), (5, 8)])
def test_flash_backward_statement_equals_autograd(N, Dh):
    torch.manual_seed(N)
    q, k, v = (torch.randn(N, Dh, dtype=torch.float64, requires_grad=True) for _ in range(3))
    do = torch.randn(N, Dh, dtype=torch.float64)
    s = (q @ k.T) / math.sqrt(Dh)
    o = torch.softmax(s, -1) @ v
    o.backward(do)
    lse2 = torch.logsumexp(s, -1).detach() * LOG2E
    dq, dk, dv = _flash_backward_statement(q.detach(), k.detach(), v.detach(), o.detach(), do, lse2)
    for name, got, ref in (("dq", dq, q.grad), ("dk", dk, k.grad), ("dv", dv, v.grad)):
        assert torch.allclose(got, ref, atol=1e-10, rtol=1e-9), f"{name}: {(got - ref).abs().max().item()}"


@pytest.mark.parametrize("N,P", [(300, 2), (1000, 4), (130, 3)])
def test_key_range_partials_merge_exactly(N, P):
    """the last-wave split and the sequence-parallel forward compute attention over P key ranges (each normalised by its own
    sum, with its log2-domain log-sum-exp) and merge: softmax over the union, no approximation"""
    torch.manual_seed(P)
    Dh = 16
    q, k, v = (torch.randn(N, Dh, dtype=torch.float64) for _ in range(3))
    c = LOG2E / math.sqrt(Dh)
    ref = torch.softmax(q @ k.T / math.sqrt(Dh), -1) @ v
    nkt = (N + 127) // 128
    tiles = (nkt + P - 1) // P          # pieces are cut at key-tile boundaries (model.cu: attention_qkv_tail_split)
    parts, lses = [], []
    for s in range(P):
        k0, k1 = s * tiles * 128, min(N, (s + 1) * tiles * 128)
        if k1 <= k0:
            break
        sc = (q @ k[k0:k1].T) * c       # log2-domain scores
        m = sc.max(-1, keepdim=True).values
        p = torch.exp2(sc - m)
        l = p.sum(-1, keepdim=True)
        parts.append((p @ v[k0:k1]) / l)                          # O_s / l_s
        lses.append((m + torch.log2(l)).squeeze(-1))              # lse_s
    lse = torch.stack(lses)                                        # [P, N]
    L = lse.max(0).values
    w = torch.exp2(lse - L)                                        # attn_merge_kernel
    out = (w[:, :, None] * torch.stack(parts)).sum(0) / w.sum(0)[:, None]
    assert torch.allclose(out, ref, atol=1e-12)
