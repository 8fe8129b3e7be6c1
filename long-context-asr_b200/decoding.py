"""Drop-in for ``lcasr/decoding/greedy.py`` (GreedyCTCDecoder): argmax -> unique_consecutive -> drop
blank, computed on the GPU (lcasr_argmax_rows + lcasr_greedy_collapse); only token ids cross PCIe."""
from __future__ import annotations

from typing import List, Union

import torch

from . import ops


class GreedyCTCDecoder(torch.nn.Module):
    """Same constructor / call contract as lcasr/decoding/greedy.py:4-22."""

    def __init__(self, tokenizer=None, blank_id=0):
        super().__init__()
        self.tokenizer = tokenizer
        self.blank = blank_id

    @torch.no_grad()
    def forward(self, emission: torch.Tensor, decode: bool = True) -> Union[str, List[int]]:
        """emission: [num_seq, num_label] log-probs / logits (CUDA, fp32)."""
        if emission.dim() != 2:
            raise ValueError(f"emission must be [num_seq, num_label], got {tuple(emission.shape)}")
        if not emission.is_cuda:
            raise RuntimeError("lcasr_b200.GreedyCTCDecoder runs on CUDA tensors only (no CPU fallback)")
        decode = decode and self.tokenizer is not None
        am = ops.argmax_rows(emission.to(torch.float32).contiguous())
        tokens, n = ops.greedy_collapse(am.view(1, -1), self.blank)
        indices = tokens[0, : int(n[0])].tolist()
        return self.tokenizer.decode(indices) if decode else indices

    @torch.no_grad()
    def decode_argmax(self, argmax: torch.Tensor) -> List[List[int]]:
        """Batched variant on the fused per-frame argmax the model already produced
        (SCConformerXL.last_argmax, int32 [B,N])."""
        tokens, n = ops.greedy_collapse(argmax.contiguous(), self.blank)
        tokens, n = tokens.cpu(), n.cpu()
        return [tokens[b, : int(n[b])].tolist() for b in range(tokens.shape[0])]
