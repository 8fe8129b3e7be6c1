"""Tensor-level wrappers over the C ABI unit operators (one per reference op on the hot path).

PyTorch is plumbing here: it owns device memory and the stream; all arithmetic runs in
liblcasr_b200.so.  Every wrapper requires CUDA tensors and raises otherwise (no CPU fallback).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib as L


def _cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise RuntimeError("lcasr_b200 ops need contiguous CUDA tensors (there is no CPU fallback)")


def _s():
    return L.current_stream()


def out_length(T: int) -> int:
    return int(L.lib.lcasr_out_length(int(T)))


def layernorm(x, weight, bias=None, eps=1e-5, kind="layer_norm", out_f32: bool = False, lo_dtype=None):
    """x fp32 [M,d] -> (fp32 output or None, low-precision/compute-dtype output or None)."""
    _cuda(x, weight, bias)
    M, d = x.shape
    o32 = torch.empty_like(x) if out_f32 else None
    olo = torch.empty(M, d, dtype=lo_dtype, device=x.device) if lo_dtype is not None else None
    L.call("lcasr_layernorm", L.ptr(x), L.ptr(weight), L.ptr(bias), M, d, float(eps),
           L.NORM_RMSNORM if kind == "rms_norm" else L.NORM_LAYERNORM, L.ptr(o32), L.ptr(olo),
           L.dtype_code(lo_dtype) if lo_dtype is not None else L.F32, _s())
    return o32, olo


def subsample_conv0(spec, w, b, out_dtype=torch.float32):
    _cuda(spec, w, b)
    B, F, T = spec.shape
    C = w.shape[0]
    out = torch.empty(B, (T - 1) // 2 + 1, (F - 1) // 2 + 1, C, dtype=out_dtype, device=spec.device)
    L.call("lcasr_subsample_conv0", L.ptr(spec), L.ptr(w), L.ptr(b), B, F, T, C, L.ptr(out), L.dtype_code(out_dtype), _s())
    return out


def subsample_dwconv(x, w, b):
    _cuda(x, w, b)
    B, Tin, Fin, C = x.shape
    out = torch.empty(B, (Tin - 1) // 2 + 1, (Fin - 1) // 2 + 1, C, dtype=x.dtype, device=x.device)
    L.call("lcasr_subsample_dwconv", L.ptr(x), L.dtype_code(x.dtype), L.ptr(w), L.ptr(b), B, Tin, Fin, C, L.ptr(out), _s())
    return out


def subsample_conv0_dw(spec, w0, b0, w1, b1):
    """fused conv0 + SiLU + first depthwise level, bf16 output [B,T2,F2,C]."""
    _cuda(spec, w0, b0, w1, b1)
    B, F, T = spec.shape
    C = w0.shape[0]
    T1, F1 = (T - 1) // 2 + 1, (F - 1) // 2 + 1
    out = torch.empty(B, (T1 - 1) // 2 + 1, (F1 - 1) // 2 + 1, C, dtype=torch.bfloat16, device=spec.device)
    L.call("lcasr_subsample_conv0_dw", L.ptr(spec), L.ptr(w0), L.ptr(b0), L.ptr(w1), L.ptr(b1), B, F, T, C, L.ptr(out), _s())
    return out


def gemm(a, w, bias=None, act=L.ACT_NONE, resid=None, alpha=1.0, out_dtype=None, impl=L.GEMM_AUTO, out=None):
    """epilogue(a[M,K] @ w[N,K]^T)."""
    _cuda(a, w, bias, resid)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.dtype == w.dtype
    if resid is not None:
        out_dtype = torch.float32
    out_dtype = out_dtype or a.dtype
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    L.call("lcasr_gemm", L.ptr(a), L.ptr(w), L.dtype_code(a.dtype), M, N, K, L.ptr(bias), act, L.ptr(resid),
           float(alpha), L.ptr(out), L.dtype_code(out_dtype), impl, _s(),
           tag=f"[{M}x{N}x{K} out={'f32' if out_dtype == torch.float32 else 'bf16'}]" if L.TIMING_TAGS else "")
    return out


def glu(x):
    _cuda(x)
    M, d2 = x.shape
    out = torch.empty(M, d2 // 2, dtype=x.dtype, device=x.device)
    L.call("lcasr_glu", L.ptr(x), L.dtype_code(x.dtype), M, d2 // 2, L.ptr(out), _s())
    return out


def rope_table(inv_freq, interp: float, N: int, offset: int = 0, transposed: bool = False):
    """cos / sin [N, Dh/2]; transposed=True: pair-major [Dh/2, N] (the layout gemm_rope reads)"""
    _cuda(inv_freq)
    half = inv_freq.numel()
    cos = torch.empty((half, N) if transposed else (N, half), dtype=torch.float32, device=inv_freq.device)
    sin = torch.empty_like(cos)
    L.call("lcasr_rope_table_t" if transposed else "lcasr_rope_table", L.ptr(inv_freq), float(interp), offset, N, half,
           L.ptr(cos), L.ptr(sin), _s())
    return cos, sin


def rope_split(qkv, B, N, H, Dh, cos=None, sin=None, v_transposed=False):
    _cuda(qkv, cos, sin)
    d = H * Dh
    q = torch.empty(B, N, H, Dh, dtype=qkv.dtype, device=qkv.device)
    k = torch.empty_like(q)
    Npad = (N + 127) // 128 * 128
    v = (torch.zeros(B, H, Dh, Npad, dtype=qkv.dtype, device=qkv.device) if v_transposed else torch.empty_like(q))
    L.call("lcasr_rope_split", L.ptr(qkv), L.dtype_code(qkv.dtype), B, N, H, Dh, L.ptr(cos), L.ptr(sin), L.ptr(q),
           L.ptr(k), L.ptr(v), int(v_transposed), Npad, _s())
    return q, k, v


def attention(q, k, v, v_transposed=False, impl=L.ATTN_AUTO):
    _cuda(q, k, v)
    B, N, H, Dh = q.shape
    Npad = v.shape[-1] if v_transposed else 0
    out = torch.empty(B, N, H * Dh, dtype=q.dtype, device=q.device)
    L.call("lcasr_attention", L.ptr(q), L.ptr(k), L.ptr(v), L.dtype_code(q.dtype), B, N, H, Dh, int(v_transposed), Npad,
           L.ptr(out), impl, _s())
    return out


def attention_window(q, k, v, left: int, right: int, kv_len=None, impl=L.ATTN_AUTO):
    """local self-attention: query i attends to keys [i - left, i + right] (-1 = unlimited)"""
    _cuda(q, k, v, kv_len)
    B, N, H, Dh = q.shape
    out = torch.empty(B, N, H * Dh, dtype=q.dtype, device=q.device)
    L.call("lcasr_attention_window", L.ptr(q), L.ptr(k), L.ptr(v), L.dtype_code(q.dtype), B, N, L.ptr(kv_len), H, Dh, int(left),
           int(right), L.ptr(out), impl, _s())
    return out


def attention_cross(q, k, v, impl=L.ATTN_AUTO):
    """q [B,Nq,H,Dh] against k,v [B,Nk,H,Dh] (sequence-parallel attention: local queries, gathered K/V)."""
    _cuda(q, k, v)
    B, Nq, H, Dh = q.shape
    Nk = k.shape[1]
    out = torch.empty(B, Nq, H * Dh, dtype=q.dtype, device=q.device)
    L.call("lcasr_attention_cross", L.ptr(q), L.ptr(k), L.ptr(v), L.dtype_code(q.dtype), B, Nq, Nk, H, Dh, L.ptr(out), impl, _s())
    return out


def dwconv_brn_silu(x, w, b, mean, std, bw, bb):
    _cuda(x, w, b, mean, std, bw, bb)
    B, N, d = x.shape
    out = torch.empty_like(x)
    L.call("lcasr_dwconv_brn_silu", L.ptr(x), L.dtype_code(x.dtype), B, N, d, w.shape[-1], L.ptr(w), L.ptr(b), L.ptr(mean),
           L.ptr(std), L.ptr(bw), L.ptr(bb), L.ptr(out), L.dtype_code(x.dtype), _s())
    return out


def softmax(x):
    _cuda(x)
    M, V = x.shape
    out = torch.empty_like(x)
    L.call("lcasr_softmax", L.ptr(x), L.dtype_code(x.dtype), M, V, L.ptr(out), L.dtype_code(x.dtype), _s())
    return out


def log_softmax_argmax_(logits) -> torch.Tensor:
    """in place on fp32 logits [M,V]; returns int32 argmax [M]."""
    _cuda(logits)
    M, V = logits.shape
    am = torch.empty(M, dtype=torch.int32, device=logits.device)
    L.call("lcasr_log_softmax_argmax", L.ptr(logits), M, V, L.ptr(am), _s())
    return am


def argmax_rows(x) -> torch.Tensor:
    _cuda(x)
    M, V = x.shape
    am = torch.empty(M, dtype=torch.int32, device=x.device)
    L.call("lcasr_argmax_rows", L.ptr(x), M, V, L.ptr(am), _s())
    return am


def greedy_collapse(argmax, blank: int, lengths=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """argmax int32 [B,N] -> (tokens int32 [B,N] (compacted prefix), n_tokens int32 [B])."""
    _cuda(argmax, lengths)
    B, N = argmax.shape
    tokens = torch.empty_like(argmax)
    n = torch.empty(B, dtype=torch.int32, device=argmax.device)
    L.call("lcasr_greedy_collapse", L.ptr(argmax), B, N, L.ptr(lengths), int(blank), L.ptr(tokens), L.ptr(n), _s())
    return tokens, n


_CTC_WS = {}


def ctc_workspace(B: int, N: int, S: int, both: bool, device):
    """hand-off buffers of the wavefront CTC recursion (cached per device; None when the wavefront form does not apply)"""
    nbytes = int(L.lib.lcasr_ctc_workspace_bytes(B, N, S, int(both)))
    if nbytes <= 0:
        return None
    ws = _CTC_WS.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _CTC_WS[device] = ws
    return ws


def ctc_wavefront_applies(B: int, N: int, S: int, both: bool) -> bool:
    return int(L.lib.lcasr_ctc_workspace_bytes(B, N, S, int(both))) > 0


def ctc_loss_fwd(log_probs, targets, input_lengths, target_lengths, blank: int, keep_alpha: bool = False):
    """log_probs fp32 [B,N,V] (batch-major); returns (nll fp32 [B], alpha or None)."""
    _cuda(log_probs, targets, input_lengths, target_lengths)
    B, N, V = log_probs.shape
    S = targets.shape[1]
    nll = torch.empty(B, dtype=torch.float32, device=log_probs.device)
    alpha = torch.empty(B, N, 2 * S + 1, dtype=torch.float32, device=log_probs.device) if keep_alpha else None
    ws = ctc_workspace(B, N, S, False, log_probs.device)
    L.call("lcasr_ctc_loss_fwd_ws", L.ptr(log_probs), B, N, V, L.ptr(targets), S, L.ptr(input_lengths), L.ptr(target_lengths),
           int(blank), L.ptr(nll), L.ptr(alpha), L.ptr(ws), 0 if ws is None else ws.numel(), _s())
    return nll, alpha


def ctc_loss_bwd(log_probs, targets, input_lengths, target_lengths, blank: int, nll, grad_nll, alpha):
    _cuda(log_probs, targets, input_lengths, target_lengths, nll, grad_nll, alpha)
    B, N, V = log_probs.shape
    S = targets.shape[1]
    beta = torch.empty_like(alpha)
    grad = torch.empty_like(log_probs)
    ws = ctc_workspace(B, N, S, False, log_probs.device)
    L.call("lcasr_ctc_loss_bwd_ws", L.ptr(log_probs), B, N, V, L.ptr(targets), S, L.ptr(input_lengths), L.ptr(target_lengths),
           int(blank), L.ptr(nll), L.ptr(grad_nll), L.ptr(alpha), L.ptr(beta), L.ptr(grad), L.ptr(ws), 0 if ws is None else ws.numel(),
           _s())
    return grad


def ctc_loss_fwd_ab(log_probs, targets, input_lengths, target_lengths, blank: int):
    """alpha and beta recursions concurrently (training): returns (nll [B], alpha, beta [B,N,2S+1])."""
    _cuda(log_probs, targets, input_lengths, target_lengths)
    B, N, V = log_probs.shape
    S = targets.shape[1]
    nll = torch.empty(B, dtype=torch.float32, device=log_probs.device)
    alpha = torch.empty(B, N, 2 * S + 1, dtype=torch.float32, device=log_probs.device)
    beta = torch.empty_like(alpha)
    ws = ctc_workspace(B, N, S, True, log_probs.device)
    L.call("lcasr_ctc_loss_fwd_ab_ws", L.ptr(log_probs), B, N, V, L.ptr(targets), S, L.ptr(input_lengths), L.ptr(target_lengths),
           int(blank), L.ptr(nll), L.ptr(alpha), L.ptr(beta), L.ptr(ws), 0 if ws is None else ws.numel(), _s())
    return nll, alpha, beta


def ctc_loss_grad(log_probs, targets, input_lengths, target_lengths, blank: int, nll, grad_nll, alpha, beta):
    _cuda(log_probs, targets, input_lengths, target_lengths, nll, grad_nll, alpha, beta)
    B, N, V = log_probs.shape
    S = targets.shape[1]
    grad = torch.empty_like(log_probs)
    L.call("lcasr_ctc_loss_grad", L.ptr(log_probs), B, N, V, L.ptr(targets), S, L.ptr(input_lengths), L.ptr(target_lengths),
           int(blank), L.ptr(nll), L.ptr(grad_nll), L.ptr(alpha), L.ptr(beta), L.ptr(grad), _s())
    return grad


def attention_partial(q, k, v):
    """one (query block x key block) term: (normalised fp32 output [B,Nq,H*Dh], log2-domain log-sum-exp [B,H,Nq])"""
    _cuda(q, k, v)
    B, Nq, H, Dh = q.shape
    out = torch.empty(B, Nq, H * Dh, dtype=torch.float32, device=q.device)
    lse = torch.empty(B, H, Nq, dtype=torch.float32, device=q.device)
    L.call("lcasr_attention_partial", L.ptr(q), L.ptr(k), L.ptr(v), B, Nq, k.shape[1], H, Dh, L.ptr(out), L.ptr(lse), _s())
    return out, lse


def attention_merge(parts, lses, H: int, Dh: int, out_dtype=torch.bfloat16):
    """parts [P, rows, H*Dh] fp32, lses [P, H, rows] fp32 -> softmax over the union of the key blocks, [rows, H*Dh]"""
    _cuda(parts, lses)
    P, rows, d = parts.shape
    out = torch.empty(rows, d, dtype=out_dtype, device=parts.device)
    L.call("lcasr_attention_merge", L.ptr(parts), L.ptr(lses), P, rows, H, Dh, L.ptr(out), L.dtype_code(out_dtype), _s())
    return out


def gemm_rope(a, w_il, cos, sin, rope_n: int, rope_cols: int, Dh: int):
    """qkv projection with the rotary embedding in the GEMM epilogue (w_il: interleaved q / k head rows; cos / sin:
    pair-major tables from rope_table(..., transposed=True)); bf16 [M, N]"""
    _cuda(a, w_il, cos, sin)
    M, K = a.shape
    N = w_il.shape[0]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    L.call("lcasr_gemm_rope", L.ptr(a), L.ptr(w_il), M, N, K, L.ptr(cos), L.ptr(sin), int(rope_n), int(rope_cols), int(Dh), L.ptr(out), _s())
    return out


def gemm_glu(a, w_glu, b_glu):
    """pointwise_conv1 + GLU in one kernel (w_glu / b_glu: 64-row blocks [32 value | 32 gate channels]); bf16 [M, N/2]"""
    _cuda(a, w_glu, b_glu)
    M, K = a.shape
    N = w_glu.shape[0]
    out = torch.empty(M, N // 2, dtype=torch.bfloat16, device=a.device)
    L.call("lcasr_gemm_glu", L.ptr(a), L.ptr(w_glu), M, N, K, L.ptr(b_glu), L.ptr(out), _s())
    return out


def attention_qkv(qkv, B: int, N: int, H: int, Dh: int, kv_len=None):
    """attention reading q, k, v as the column blocks of the [B*N, 3*H*Dh] projection (no split pass); bf16 [B, N, H*Dh]"""
    _cuda(qkv, kv_len)
    out = torch.empty(B, N, H * Dh, dtype=torch.bfloat16, device=qkv.device)
    L.call("lcasr_attention_qkv", L.ptr(qkv), B, N, L.ptr(kv_len), H, Dh, L.ptr(out), _s())
    return out


def layernorm_chain(x, weights, biases, eps=1e-5, kind="layer_norm", f32_stage=None, lo_dtype=None):
    """up to three norms back to back on rows kept in registers: (fp32 result of stage `f32_stage` or None, low-precision
    result of the last stage or None)"""
    import ctypes as C
    _cuda(x, *weights, *[b for b in biases if b is not None])
    M, d = x.shape
    n = len(weights)
    o32 = torch.empty_like(x) if f32_stage is not None else None
    olo = torch.empty(M, d, dtype=lo_dtype, device=x.device) if lo_dtype is not None else None
    wp = (C.c_void_p * n)(*[w.data_ptr() for w in weights])
    bp = (C.c_void_p * n)(*[(b.data_ptr() if b is not None else None) for b in biases])
    L.call("lcasr_layernorm_chain", L.ptr(x), n, wp, bp, M, d, float(eps),
           L.NORM_RMSNORM if kind == "rms_norm" else L.NORM_LAYERNORM, 0 if f32_stage is None else int(f32_stage), L.ptr(o32),
           L.ptr(olo), L.dtype_code(lo_dtype) if lo_dtype is not None else L.F32, _s())
    return o32, olo
