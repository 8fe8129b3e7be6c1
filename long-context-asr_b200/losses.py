"""Drop-in for the reference's CTC loss call sites (exp/train.py:104,249; lcasr/eval/dynamic_eval.py:47):
``CTCLoss(blank=V, reduction='sum')(log_probs.transpose(0,1), targets, input_lengths, target_lengths)``.
Forward = alpha recursion, backward = beta recursion + class scatter, both in liblcasr_b200.so."""
from __future__ import annotations

import torch

from . import ops


class _CTCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lp_bnv, targets, input_lengths, target_lengths, blank):
        need_grad = lp_bnv.requires_grad
        B_, N_ = lp_bnv.shape[0], lp_bnv.shape[1]
        # alpha and beta in one launch: up to 4096 extended states on one SM each, or any length as two wavefronts
        ctx.concurrent = need_grad and (2 * targets.shape[1] + 1 <= 4096 or ops.ctc_wavefront_applies(B_, N_, targets.shape[1], True))
        if ctx.concurrent:  # alpha and beta in one launch: the backward is only the class scatter
            nll, alpha, beta = ops.ctc_loss_fwd_ab(lp_bnv, targets, input_lengths, target_lengths, blank)
            ctx.save_for_backward(lp_bnv, targets, input_lengths, target_lengths, nll, alpha, beta)
            ctx.blank = blank
            return nll
        nll, alpha = ops.ctc_loss_fwd(lp_bnv, targets, input_lengths, target_lengths, blank, keep_alpha=need_grad)
        if need_grad:
            ctx.save_for_backward(lp_bnv, targets, input_lengths, target_lengths, nll, alpha)
            ctx.blank = blank
        return nll

    @staticmethod
    def backward(ctx, grad_nll):
        if ctx.concurrent:
            lp_bnv, targets, input_lengths, target_lengths, nll, alpha, beta = ctx.saved_tensors
            grad = ops.ctc_loss_grad(lp_bnv, targets, input_lengths, target_lengths, ctx.blank, nll,
                                     grad_nll.to(torch.float32).contiguous(), alpha, beta)
            return grad, None, None, None, None
        lp_bnv, targets, input_lengths, target_lengths, nll, alpha = ctx.saved_tensors
        grad = ops.ctc_loss_bwd(lp_bnv, targets, input_lengths, target_lengths, ctx.blank, nll,
                                grad_nll.to(torch.float32).contiguous(), alpha)
        return grad, None, None, None, None


class CTCLoss(torch.nn.Module):
    """torch.nn.CTCLoss-compatible (log_probs [T,B,C] time-major as the reference passes it)."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        if reduction not in ("none", "mean", "sum"):
            raise ValueError(f"bad reduction {reduction!r}")
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        if not log_probs.is_cuda:
            raise RuntimeError("lcasr_b200.CTCLoss runs on CUDA tensors only (no CPU fallback)")
        if log_probs.dim() != 3:
            raise ValueError("log_probs must be [T, B, C]")
        if targets.dim() != 2:
            raise NotImplementedError("1-D concatenated targets are not used by the reference call sites")
        dev = log_probs.device
        # [T,B,C] -> batch-major; free when the caller passed model_out.transpose(0,1) as the reference does
        lp = log_probs.transpose(0, 1)
        if lp.dtype != torch.float32 or not lp.is_contiguous():
            lp = lp.to(torch.float32).contiguous()
        B, N, _ = lp.shape
        tg = torch.as_tensor(targets).to(device=dev, dtype=torch.int64).contiguous()
        il = torch.as_tensor(input_lengths).to(device=dev, dtype=torch.int32).contiguous()
        tl = torch.as_tensor(target_lengths).to(device=dev, dtype=torch.int64).contiguous()
        if tg.shape[1] == 0:
            tg = torch.zeros(B, 1, dtype=torch.int64, device=dev)
        nll = _CTCFunction.apply(lp, tg, il, tl, self.blank)
        if self.zero_infinity:
            nll = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
        if self.reduction == "none":
            return nll
        if self.reduction == "sum":
            return nll.sum()
        return (nll / tl.clamp_min(1).to(nll.dtype)).mean()
