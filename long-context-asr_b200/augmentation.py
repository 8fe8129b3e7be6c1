"""SpecAugment on the GPU — drop-in for ``lcasr/utils/augmentation.py:10-104`` (the module ``exp/train.py:227`` applies to
the spectrogram batch right before the encoder).

Same constructor keywords, same call ``SpecAugment(...)(specgram, audio_lengths)`` and the same consumption of the torch
random stream as the reference (per mask: ``torch.rand`` for the width, then for the position; on the spectrogram's
device with shape ``specgram.shape[:-2]`` for iid masks, ``torch.rand(1)`` on the host otherwise), so a run seeded like
the reference masks the same cells.  The arithmetic is two kernels (``csrc/augment.cu``): the mean over the un-padded
frames (device scalar, no host sync) and ONE pass that applies all time and frequency masks — the reference makes
n_time + n_freq full passes.  No CPU path."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L


def _limited(mask_param: int, p: float, axis_length: int) -> int:
    # torchaudio.functional._get_mask_param
    return mask_param if p == 1.0 else min(mask_param, int(axis_length * p))


class SpecAugment(torch.nn.Module):
    def __init__(self, n_time_masks: int, n_freq_masks: int, freq_mask_param: int, iid_masks: bool = True,
                 time_mask_param: int = -1, min_p: float = -1, max_p: float = 1.0, zero_masking: bool = False, **kwargs) -> None:
        super().__init__()
        # the same argument checks as the reference constructor (augmentation.py:49-51)
        assert n_time_masks == 0 or min_p != -1 or time_mask_param != -1, "time masks need a width: give time_mask_param or min_p"
        assert min_p == -1 or 0 <= min_p <= 1, f"min_p={min_p} is not a proportion in [0, 1]"
        assert 0 <= max_p <= 1, f"max_p={max_p} is not a proportion in [0, 1]"
        self.n_time_masks, self.n_freq_masks = n_time_masks, n_freq_masks
        self.time_mask_param, self.freq_mask_param = time_mask_param, freq_mask_param
        self.min_p, self.max_p = min_p, max_p
        self.iid_masks, self.zero_masking = iid_masks, zero_masking

    def mask_params(self, f: int, t: int):
        """effective (time, freq) mask parameters for a [.., f, t] input (augmentation.py:79-82 + torchaudio's max_p limit)"""
        tw = self.time_mask_param
        if self.min_p != -1:
            tw = int(int(t * self.min_p) / self.n_time_masks) if self.n_time_masks != 0 else 0
        return _limited(tw, self.max_p, t), _limited(self.freq_mask_param, self.max_p, f)

    def draw(self, specgram: torch.Tensor):
        """the uniform draws, consumed exactly like the reference does: returns (time_param, u_time, freq_param, u_freq)
        with u_* of shape [n_masks, 2, draws_per_mask] on the spectrogram's device (None when that axis is off)"""
        f, t = specgram.shape[-2:]
        tp, fp = self.mask_params(f, t)
        iid = specgram.dim() > 2 and self.iid_masks is True
        lead = specgram.shape[:-2] if specgram.dim() > 3 else (specgram.shape[0], 1) if specgram.dim() == 3 else ()

        def draws(n):
            rows = []
            for _ in range(n):
                if iid:
                    rows.append(torch.stack([torch.rand(lead, device=specgram.device, dtype=specgram.dtype).reshape(-1),
                                             torch.rand(lead, device=specgram.device, dtype=specgram.dtype).reshape(-1)]))
                else:
                    rows.append(torch.stack([torch.rand(1), torch.rand(1)]))
            return torch.stack(rows).to(specgram.device, torch.float32).contiguous()

        u_time = draws(self.n_time_masks) if (tp >= 1 and self.n_time_masks > 0) else None
        u_freq = draws(self.n_freq_masks) if (fp >= 1 and self.n_freq_masks > 0) else None
        return tp, u_time, fp, u_freq

    def forward(self, specgram: torch.Tensor, audio_lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        tp, u_time, fp, u_freq = self.draw(specgram)
        return apply_masks(specgram, audio_lengths, tp, u_time, fp, u_freq, self.zero_masking)


def apply_masks(specgram: torch.Tensor, audio_lengths, time_param: int, u_time, freq_param: int, u_freq,
                zero_masking: bool = False) -> torch.Tensor:
    """out = fill where a time or frequency mask covers the cell, specgram elsewhere; u_* as returned by
    ``SpecAugment.draw`` ([n_masks, 2, B] iid or [n_masks, 2, 1] shared)."""
    if not specgram.is_cuda:
        raise RuntimeError("lcasr_b200.SpecAugment runs on CUDA tensors only (no CPU fallback)")
    if specgram.dtype != torch.float32:
        raise TypeError("SpecAugment expects the fp32 spectrogram the reference feeds it (exp/train.py:224)")
    shape = specgram.shape
    f, t = shape[-2:]
    x = specgram.contiguous().view(-1, f, t)
    B = x.shape[0]
    acc = None
    if not zero_masking:
        lens = None
        if audio_lengths is not None:
            assert audio_lengths.numel() == B, "audio_lengths: one length per spectrogram"
            lens = audio_lengths.to(x.device, torch.int32).contiguous()
        acc = torch.empty(2, dtype=torch.float64, device=x.device)
        L.call("lcasr_specaug_mean", L.ptr(x), B, f, t, L.ptr(lens), L.ptr(acc), L.current_stream())
    n_time = 0 if u_time is None else u_time.shape[0]
    n_freq = 0 if u_freq is None else u_freq.shape[0]
    if n_time + n_freq == 0:
        return specgram
    per = (u_time if u_time is not None else u_freq).shape[-1]
    out = torch.empty_like(x)
    L.call("lcasr_specaug_apply", L.ptr(x), B, f, t, n_time, max(time_param, 1), L.ptr(u_time), n_freq, max(freq_param, 1),
           L.ptr(u_freq), per, L.ptr(acc), L.ptr(out), L.current_stream())
    return out.view(shape)
