"""GPU front-end — drop-in for ``lcasr/utils/audio_tools.py:44-57`` (``to_spectogram``): waveform -> 80-bin mel power
spectrogram (torchaudio ``MelSpectrogram`` with the reference's settings) -> per-bin standardisation over time.  The
window-premultiplied DFT twiddles and the HTK mel filterbank are built once on the host from their definitions (no
torchaudio in the product path) and cached on the device; the arithmetic is ``lcasr_melspec`` (csrc/frontend.cu)."""
from __future__ import annotations

import math

import torch

from . import _lib as L

WIN_LENGTH, HOP_LENGTH, SR, N_FFT, N_MELS = 400, 160, 16000, 512, 80
_TABLES = {}


def _hz_to_mel(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def build_tables(n_mels: int = N_MELS):
    """(cos_tab, sin_tab [512,257], fb [257,n_mels]) in float64 on the CPU, from the definitions:
    torch.hann_window(400) (periodic) centred in the 512-sample frame (torch.stft pads the window to n_fft),
    X[k] = sum_n w[n] x[n] e^{-2 pi i k n / 512}; torchaudio.functional.melscale_fbanks(257, 0, 8000, n_mels, 16000, norm=None,
    mel_scale='htk')."""
    n = torch.arange(N_FFT, dtype=torch.float64)
    win = torch.zeros(N_FFT, dtype=torch.float64)
    left = (N_FFT - WIN_LENGTH) // 2
    win[left:left + WIN_LENGTH] = 0.5 - 0.5 * torch.cos(2.0 * math.pi * torch.arange(WIN_LENGTH, dtype=torch.float64) / WIN_LENGTH)
    k = torch.arange(N_FFT // 2 + 1, dtype=torch.float64)
    ang = 2.0 * math.pi * torch.outer(n, k) / N_FFT
    cos_tab, sin_tab = win[:, None] * torch.cos(ang), -win[:, None] * torch.sin(ang)
    all_freqs = torch.linspace(0, SR // 2, N_FFT // 2 + 1, dtype=torch.float64)
    m_pts = torch.linspace(_hz_to_mel(0.0), _hz_to_mel(SR / 2.0), n_mels + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.minimum(down, up), min=0.0)
    return cos_tab, sin_tab, fb


def _device_tables(device, n_mels):
    key = (str(device), n_mels)
    if key not in _TABLES:
        _TABLES[key] = tuple(t.to(torch.float32).contiguous().to(device) for t in build_tables(n_mels))
    return _TABLES[key]


@torch.no_grad()
def to_spectogram(waveform: torch.Tensor, global_normalisation: bool = True, n_mels: int = N_MELS) -> torch.Tensor:
    """waveform [C, n_samples] (or [n_samples]) fp32 at 16 kHz, CUDA -> [C, n_mels, 1 + n_samples // 160] fp32"""
    if not waveform.is_cuda:
        raise RuntimeError("lcasr_b200.frontend runs on CUDA tensors only (there is no CPU fallback)")
    w = waveform.to(torch.float32)
    if w.dim() == 1:
        w = w[None]
    w = w.contiguous()
    B, n = w.shape
    cos_tab, sin_tab, fb = _device_tables(w.device, n_mels)
    frames = int(L.lib.lcasr_melspec_frames(n))
    out = torch.empty(B, n_mels, frames, dtype=torch.float32, device=w.device)
    sums = torch.empty(B, n_mels, 2, dtype=torch.float64, device=w.device)
    L.call("lcasr_melspec", L.ptr(w), B, n, L.ptr(cos_tab), L.ptr(sin_tab), L.ptr(fb), n_mels, L.ptr(out), L.ptr(sums),
           int(bool(global_normalisation)), L.current_stream())
    return out


def total_seconds(spectogram_length: int) -> float:
    return (spectogram_length * HOP_LENGTH) / SR


def total_frames(seconds: float) -> int:
    return int((seconds * 16000) / HOP_LENGTH)
