"""Tensor-level wrappers over the TRAINING entry points of the C ABI (include/lcasr_b200.h, "Training step").

PyTorch owns memory and the stream; every arithmetic operation of the backward pass is one of the
hand-written kernels in csrc/{gemm_tcx,train,subsample_bwd}.cu.  bf16 activations, fp32 gradients of
parameters and of the residual stream.  No CPU fallback.
"""
from __future__ import annotations

import os

import ctypes as C

import torch

from . import _lib as L
from .ops import _cuda, _s

BF = torch.bfloat16


def gemm_ex(A, B, out, M, N, K, *, a_mn=False, b_mn=False, lda, ldb, ldo, nb=(1, 1), sa=(0, 0), sb=(0, 0), so=(0, 0),
            aux=None, ldaux=0, sx=(0, 0), rowvec=None, sr=(0, 0), alpha=1.0, epi=L.EPI_SCALE, ksplit=0):
    """acc[M,N] = sum_k A(m,k) B(n,k) per batch entry; see lcasr_gemm_ex in the header for the layouts."""
    args = L.LcasrGemmExArgs(
        A=L.ptr(A), B=L.ptr(B), out=L.ptr(out), aux=L.ptr(aux), rowvec=L.ptr(rowvec), M=M, N=N, K=K, a_mn=int(a_mn),
        b_mn=int(b_mn), lda=lda, ldb=ldb, ldo=ldo, ldaux=ldaux, nb1=nb[0], nb2=nb[1], sa1=sa[0], sa2=sa[1], sb1=sb[0],
        sb2=sb[1], so1=so[0], so2=so[1], sx1=sx[0], sx2=sx[1], sr1=sr[0], sr2=sr[1], alpha=float(alpha), epi=epi,
        out_dtype=L.dtype_code(out.dtype), ksplit=ksplit)
    L.call("lcasr_gemm_ex", C.byref(args), _s(),
           tag=f"[{M}x{N}x{K} a_mn={int(a_mn)} b_mn={int(b_mn)} out={'f32' if out.dtype == torch.float32 else 'bf16'} epi={epi} "
               f"batch={nb[0] * nb[1]}]" if L.TIMING_TAGS else "")
    return out


def dgrad(dy, w, out=None, aux=None, epi=L.EPI_SCALE, alpha=1.0):
    """dx[M,in] = alpha * dy[M,out] @ w[out,in]  (bf16 out, or fp32 `out` that is accumulated into)."""
    _cuda(dy, w, aux)
    M, No = dy.shape
    Ni = w.shape[1]
    assert w.shape[0] == No
    if out is None:
        out = torch.empty(M, Ni, dtype=BF, device=dy.device)
    return gemm_ex(dy, w, out, M, Ni, No, b_mn=True, lda=No, ldb=Ni, ldo=Ni, aux=aux, ldaux=Ni, epi=epi, alpha=alpha)


def wgrad(dy, x, dw, alpha=1.0):
    """dw[out,in] (fp32) += alpha * dy[M,out]^T @ x[M,in]; split-K over the tokens."""
    _cuda(dy, x, dw)
    M, No = dy.shape
    Ni = x.shape[1]
    assert x.shape[0] == M and tuple(dw.shape) == (No, Ni) and dw.dtype == torch.float32
    return gemm_ex(dy, x, dw, No, Ni, M, a_mn=True, b_mn=True, lda=No, ldb=Ni, ldo=Ni, alpha=alpha)


def attention_train(q, k, v, kv_len=None):
    """q,k,v bf16 [B,N,H,Dh] -> (out [B,N,H*Dh] bf16, lse2 [B,H,N] fp32 in the log2 domain).
    kv_len int32 [B] (device): keys at or beyond it are masked (padded batch)."""
    _cuda(q, k, v)
    B, N, H, Dh = q.shape
    out = torch.empty(B, N, H * Dh, dtype=BF, device=q.device)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=q.device)
    if kv_len is None:
        L.call("lcasr_attention_train", L.ptr(q), L.ptr(k), L.ptr(v), B, N, H, Dh, L.ptr(out), L.ptr(lse), _s())
    else:
        L.call("lcasr_attention_train_masked", L.ptr(q), L.ptr(k), L.ptr(v), B, N, L.ptr(kv_len), H, Dh, L.ptr(out), L.ptr(lse), _s())
    return out, lse


def glu_masked(u, lengths, B: int, N: int):
    """GLU of u [B*N, 2d] with the rows of padded tokens written as zeros (convolution.py:107-110)"""
    _cuda(u, lengths)
    M, d2 = u.shape
    out = torch.empty(M, d2 // 2, dtype=u.dtype, device=u.device)
    L.call("lcasr_glu_masked", L.ptr(u), L.dtype_code(u.dtype), B, N, d2 // 2, L.ptr(lengths), L.ptr(out), _s())
    return out


def mask_rows_(x, lengths, B: int, N: int):
    """in place: zero the rows n >= lengths[b] of x viewed as [B, N, -1] (lengths int32 [B] on the device)"""
    _cuda(x, lengths)
    L.call("lcasr_mask_rows", L.ptr(x), L.dtype_code(x.dtype), B, N, x.numel() // (B * N), L.ptr(lengths), _s())
    return x


def flash_applies(Dh: int) -> bool:
    """the flash-style backward kernel exists for head dims 64 and 128; LCASR_ATTN_BWD_FLASH=0 selects the materialised form"""
    return Dh in (64, 128) and os.environ.get("LCASR_ATTN_BWD_FLASH", "1") != "0"


def attention_bwd(q, k, v, o, do, lse, chunk_b: int = 0, fused_pds: bool = True, lens=None, flash=None):
    """Backward of softmax(q k^T / sqrt(Dh)) v from the saved output and log-sum-exp.
    q,k,v,o,do bf16 [B,N,H,Dh]; returns dq, dk, dv (same layout).  Head dims 64 / 128 (flash=None: automatic): the
    flash-style kernels (lcasr_attention_bwd_flash: O(N) memory).  Otherwise P and dS are materialised per group of
    recordings ([b,H,N,N] bf16 each, sized to stay L2-resident) and every product is one batched tcgen05 GEMM.
    `lens` (host ints, tokens per recording): padded batch — recording b is differentiated as the n_b x n_b problem of
    its valid tokens (masked keys have P = 0 and the output rows of padded queries are zeroed by the caller, so their
    dO is zero: nothing else contributes); gradient rows of padded tokens are zero."""
    _cuda(q, k, v, o, do, lse)
    B, N, H, Dh = q.shape
    d = H * Dh
    dev = q.device
    ragged = lens is not None and any(int(n) != N for n in lens)
    alloc = torch.zeros_like if ragged else torch.empty_like
    dq, dk, dv = alloc(q), alloc(q), alloc(q)
    if flash_applies(Dh) if flash is None else flash:
        # flash-style backward: no [N, N] tensor reaches memory (two launches: dk / dv, then dq)
        groups = [(b, 1, int(lens[b])) for b in range(B)] if ragged else [(0, B, N)]
        for b0, nb, n in groups:
            if n <= 0:
                continue
            ws_bytes = int(L.lib.lcasr_attention_bwd_flash_workspace_bytes(nb, n, H))
            ws = torch.empty(ws_bytes // 8, 2, dtype=torch.float32, device=dev)
            L.call("lcasr_attention_bwd_flash", L.ptr(q[b0:]), L.ptr(k[b0:]), L.ptr(v[b0:]), L.ptr(o[b0:]), L.ptr(do[b0:]),
                   L.ptr(lse[b0:]), nb, n, N, N, H, Dh, L.ptr(dq[b0:]), L.ptr(dk[b0:]), L.ptr(dv[b0:]), L.ptr(ws), ws_bytes, _s())
        return dq, dk, dv
    Dvec = torch.empty(B, H, N, dtype=torch.float32, device=dev)
    L.call("lcasr_rowdot", L.ptr(do), L.ptr(o), B, N, H, Dh, L.ptr(Dvec), _s())
    if chunk_b <= 0:  # P + dS of one chunk ~ 96 MB
        chunk_b = max(1, min(B, (48 << 20) // (H * N * N * 2)))
    if ragged:
        chunk_b = 1
    Npad = (N + 7) // 8 * 8
    need = 2 * chunk_b * H * N * Npad * 2  # P and dS, bf16
    free, _ = torch.cuda.mem_get_info(dev)
    free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)  # blocks the caching allocator can reuse
    if need > free:
        raise RuntimeError(
            f"attention backward (materialised form, head_dim {Dh}): P and dS of one recording need {need / 2**30:.1f} GiB "
            f"([{H}, {N}, {N}] bf16 each) but {free / 2**30:.1f} GiB are free; the flash-style backward (O(N) memory) covers "
            "head dims 64 and 128 only")
    P = torch.empty(chunk_b * H * N * Npad, dtype=BF, device=dev)
    dS = torch.empty_like(P)
    scale = 1.0 / (Dh ** 0.5)
    for b0 in range(0, B, chunk_b):
        nb = min(chunk_b, B - b0)
        n = int(lens[b0]) if ragged else N   # valid tokens of this chunk's recordings
        Np = (n + 7) // 8 * 8               # row pitch of the score matrices
        sNN = n * Np
        qs, ks, vs, dos = q[b0:], k[b0:], v[b0:], do[b0:]
        lse_c, dvec_c = lse[b0:], Dvec[b0:]
        r4 = (N, H * N)
        if n != N:  # the fused kernel derives the row-vector strides from n
            lse_c, dvec_c = lse[b0, :, :n].contiguous(), Dvec[b0, :, :n].contiguous()
            r4 = (n, H * n)
        bat = dict(nb=(H, nb))
        x4 = (Dh, N * d)  # (head, recording) strides of a [B,N,H,Dh] tensor
        s4 = (sNN, H * sNN)
        if fused_pds:  # P = exp2(scale*log2e * q k^T - lse2) and dS = P o (do v^T - D) * scale in one pass
            L.call("lcasr_attention_bwd_pds", L.ptr(qs), L.ptr(ks), L.ptr(vs), L.ptr(dos), L.ptr(lse_c), L.ptr(dvec_c),
                   nb, n, H, Dh, L.ptr(P), L.ptr(dS), _s())
        else:          # the same as two GEMMs with epilogues (kept as the A/B reference of the fused kernel)
            gemm_ex(qs, ks, P, n, n, Dh, lda=d, ldb=d, ldo=Np, sa=x4, sb=x4, so=s4, rowvec=lse_c, sr=r4,
                    alpha=scale * 1.4426950408889634, epi=L.EPI_EXP2, **bat)
            gemm_ex(dos, vs, dS, n, n, Dh, lda=d, ldb=d, ldo=Np, sa=x4, sb=x4, so=s4, aux=P, ldaux=Np, sx=s4,
                    rowvec=dvec_c, sr=r4, alpha=scale, epi=L.EPI_DS, **bat)
        # The three products are independent and each fills only H*ceil(N/128) of the 148 SMs: dk and dv run on side
        # streams next to dq (fork after P/dS are written, join before they are overwritten / the results are used).
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        # dq = dS k       (B operand k stored [keys, Dh] = [K, N]: MN-major)
        gemm_ex(dS, ks, dq[b0:], n, Dh, n, b_mn=True, lda=Np, ldb=d, ldo=d, sa=s4, sb=x4, so=x4, **bat)
        for i, side in enumerate(_side_streams(dev)):
            side.wait_event(fork)
            with torch.cuda.stream(side):
                if i == 0:  # dk = dS^T q     (A stored [q, keys] = [K, M]: MN-major)
                    gemm_ex(dS, qs, dk[b0:], n, Dh, n, a_mn=True, b_mn=True, lda=Np, ldb=d, ldo=d, sa=s4, sb=x4, so=x4, **bat)
                else:       # dv = P^T do
                    gemm_ex(P, dos, dv[b0:], n, Dh, n, a_mn=True, b_mn=True, lda=Np, ldb=d, ldo=d, sa=s4, sb=x4, so=x4, **bat)
                join = torch.cuda.Event()
                join.record(side)
            main.wait_event(join)
    return dq, dk, dv


_SIDE = {}


def _side_streams(dev):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _SIDE:
        _SIDE[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return _SIDE[key]


def gemm_act_pre(a, w, bias, act):
    """(act(pre), pre) with pre = a @ w^T + bias, both bf16: one tcgen05 GEMM with a two-output epilogue."""
    _cuda(a, w, bias)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=BF, device=a.device)
    pre = torch.empty_like(out)
    L.call("lcasr_gemm_act_pre", L.ptr(a), L.ptr(w), M, N, K, L.ptr(bias), act, L.ptr(out), L.ptr(pre), _s(),
           tag=f"[{M}x{N}x{K} out=bf16+pre]" if L.TIMING_TAGS else "")
    return out, pre


def scale_cast(x, scale=1.0):
    _cuda(x)
    out = torch.empty(x.shape, dtype=BF, device=x.device)
    L.call("lcasr_scale_cast", L.ptr(x), x.numel(), float(scale), L.ptr(out), _s())
    return out


def act_fwd(x, act):
    _cuda(x)
    out = torch.empty_like(x)
    L.call("lcasr_act_fwd", L.ptr(x), x.numel(), act, L.ptr(out), _s())
    return out


def act_bwd(pre, dy, act):
    _cuda(pre, dy)
    out = torch.empty_like(pre)
    L.call("lcasr_act_bwd", L.ptr(pre), L.ptr(dy), pre.numel(), act, L.ptr(out), _s())
    return out


def add_bf16_(x, a):
    _cuda(x, a)
    assert x.dtype == torch.float32 and a.dtype == BF and x.numel() == a.numel()
    L.call("lcasr_add_bf16", L.ptr(x), L.ptr(a), x.numel(), _s())
    return x


def glu_bwd(u, dg):
    _cuda(u, dg)
    M, d = dg.shape
    du = torch.empty_like(u)
    L.call("lcasr_glu_bwd", L.ptr(u), L.ptr(dg), M, d, L.ptr(du), _s())
    return du


def rope_bwd_merge(dq, dk, dv, cos=None, sin=None):
    _cuda(dq, dk, dv, cos, sin)
    B, N, H, Dh = dq.shape
    out = torch.empty(B * N, 3 * H * Dh, dtype=BF, device=dq.device)
    L.call("lcasr_rope_bwd_merge", L.ptr(dq), L.ptr(dk), L.ptr(dv), B, N, H, Dh, L.ptr(cos), L.ptr(sin), L.ptr(out), _s())
    return out


def softmax_bwd(p, dp, scale=1.0, out=None):
    _cuda(p, dp)
    M, V = p.shape
    out = torch.empty_like(p) if out is None else out
    L.call("lcasr_softmax_bwd", L.ptr(p), L.ptr(dp), M, V, float(scale), L.ptr(out), _s())
    return out


def log_softmax_bwd(lp, dlp, scale=1.0):
    _cuda(lp, dlp)
    M, V = lp.shape
    out = torch.empty(M, V, dtype=BF, device=lp.device)
    L.call("lcasr_log_softmax_bwd", L.ptr(lp), L.ptr(dlp), M, V, float(scale), L.ptr(out), _s())
    return out


def colsum_(out, x, scale=1.0):
    """out[d] (fp32) += scale * x.sum(0)"""
    _cuda(out, x)
    M, d = x.shape
    L.call("lcasr_colsum", L.ptr(x), L.dtype_code(x.dtype), M, d, float(scale), L.ptr(out), _s())
    return out


def layernorm_bwd(x, dy, weight, dx, dweight, dbias=None, eps=1e-5, kind="layer_norm", accumulate=True, cast_scale=None):
    """returns dx, or (dx, bf16 copy of cast_scale * dx) when cast_scale is given (the next sub-layer's dY operand)"""
    _cuda(x, dy, weight, dx, dweight, dbias)
    M, d = x.shape
    cast = torch.empty(M, d, dtype=BF, device=x.device) if cast_scale is not None else None
    L.call("lcasr_layernorm_bwd_cast", L.ptr(x), L.ptr(dy), L.dtype_code(dy.dtype), L.ptr(weight), M, d, float(eps),
           L.NORM_RMSNORM if kind == "rms_norm" else L.NORM_LAYERNORM, int(accumulate), L.ptr(dx), L.ptr(dweight),
           L.ptr(dbias), L.ptr(cast), float(cast_scale or 1.0), _s())
    return dx if cast_scale is None else (dx, cast)


def dwconv1d_fwd(x, w, b, stats=False):
    _cuda(x, w, b)
    B, N, d = x.shape
    out = torch.empty_like(x)
    s = torch.zeros(2, d, dtype=torch.float64, device=x.device) if stats else None  # fp64 cross-CTA sums (reproducible)
    L.call("lcasr_dwconv1d_fwd", L.ptr(x), B, N, d, w.shape[-1], L.ptr(w), L.ptr(b), L.ptr(out),
           L.ptr(s[0]) if stats else None, L.ptr(s[1]) if stats else None, _s())
    return out, s


def dwconv1d_bwd_data(dout, w):
    _cuda(dout, w)
    B, N, d = dout.shape
    din = torch.empty_like(dout)
    L.call("lcasr_dwconv1d_bwd_data", L.ptr(dout), B, N, d, w.shape[-1], L.ptr(w), L.ptr(din), _s())
    return din


def dwconv1d_bwd_weight_(x, dout, dw, db):
    _cuda(x, dout, dw, db)
    B, N, d = x.shape
    L.call("lcasr_dwconv1d_bwd_weight", L.ptr(x), L.ptr(dout), B, N, d, dw.shape[-1], L.ptr(dw), L.ptr(db), _s())


def brn_train_stats(sums, count, running_mean, running_std, eps, rmax, dmax, momentum, weight, bias):
    """returns (A, Bc, stats[5,d]); updates running_mean / running_std in place."""
    d = weight.numel()
    dev = weight.device
    A = torch.empty(d, dtype=torch.float32, device=dev)
    Bc = torch.empty_like(A)
    stats = torch.empty(5, d, dtype=torch.float32, device=dev)
    L.call("lcasr_brn_train_stats", L.ptr(sums[0]), L.ptr(sums[1]), int(count), d, L.ptr(running_mean), L.ptr(running_std),
           float(eps), float(rmax), float(dmax), float(momentum), L.ptr(weight), L.ptr(bias), L.ptr(A), L.ptr(Bc),
           L.ptr(stats), _s())
    return A, Bc, stats


def affine_silu(c, A, Bc):
    _cuda(c, A, Bc)
    out = torch.empty_like(c)
    L.call("lcasr_affine_silu", L.ptr(c), c.numel() // c.shape[-1], c.shape[-1], L.ptr(A), L.ptr(Bc), L.ptr(out), _s())
    return out


def brn_silu_bwd(c, dy, A, Bc, stats, weight, dweight, dbias, eval_mode: bool = False):
    """backward of y = silu(BatchRenorm(c)): returns dc (bf16); dweight, dbias (fp32) accumulate.
    eval_mode: the statistics are the running buffers (constants): dc = dz * weight / running_std, without the
    mean / variance terms of the batch-statistics form."""
    _cuda(c, dy, A, Bc, stats, weight, dweight, dbias)
    d = c.shape[-1]
    M = c.numel() // d
    dz = torch.empty_like(c)
    S = torch.zeros(2, d, dtype=torch.float64, device=c.device)  # fp64 cross-CTA sums (reproducible)
    L.call("lcasr_affine_silu_bwd", L.ptr(c), L.ptr(dy), M, d, L.ptr(A), L.ptr(Bc), L.ptr(stats), L.ptr(dz), L.ptr(S[0]),
           L.ptr(S[1]), _s())
    coef = torch.empty(3, d, dtype=torch.float32, device=c.device)
    L.call("lcasr_brn_bwd_finalize", L.ptr(S[0]), L.ptr(S[1]), M, d, L.ptr(weight), L.ptr(stats), L.ptr(dweight),
           L.ptr(dbias), L.ptr(coef), _s())
    if eval_mode:
        coef[1:].zero_()
    dc = torch.empty_like(c)
    L.call("lcasr_affine3", L.ptr(dz), L.ptr(c), M, d, L.ptr(coef), L.ptr(dc), _s())
    return dc


def subsample_dwconv_bwd_data(dout, w, Tin, Fin):
    _cuda(dout, w)
    B, C_ = dout.shape[0], dout.shape[-1]
    din = torch.empty(B, Tin, Fin, C_, dtype=BF, device=dout.device)
    L.call("lcasr_subsample_dwconv_bwd_data", L.ptr(dout), L.ptr(w), B, Tin, Fin, C_, L.ptr(din), _s())
    return din


def subsample_dwconv_bwd_weight_(x, dout, dw, db):
    _cuda(x, dout, dw, db)
    B, Tin, Fin, C_ = x.shape
    L.call("lcasr_subsample_dwconv_bwd_weight", L.ptr(x), L.ptr(dout), B, Tin, Fin, C_, L.ptr(dw), L.ptr(db), _s())


def subsample_l1_bwd_(spec, w0, b0, w1, dd1, dw0, db0, dw1, db1):
    """fused backward of conv0 + SiLU + depthwise level 1: parameter gradients (+=) straight from the spectrogram and
    dd1 [B,T2,F2,C]; the conv0 activation / its gradient are never materialised"""
    _cuda(spec, w0, b0, w1, dd1, dw0, db0, dw1, db1)
    B, F, T = spec.shape
    L.call("lcasr_subsample_l1_bwd", L.ptr(spec), L.ptr(w0), L.ptr(b0), L.ptr(w1), L.ptr(dd1), B, F, T, w0.shape[0],
           L.ptr(dw0), L.ptr(db0), L.ptr(dw1), L.ptr(db1), _s())


def subsample_conv0_bwd_(spec, w, b, ds1, dw, db):
    _cuda(spec, w, b, ds1, dw, db)
    B, F, T = spec.shape
    L.call("lcasr_subsample_conv0_bwd", L.ptr(spec), L.ptr(w), L.ptr(b), L.ptr(ds1), B, F, T, w.shape[0], L.ptr(dw),
           L.ptr(db), _s())
