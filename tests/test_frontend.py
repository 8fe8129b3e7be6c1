"""Front-end (lcasr/utils/audio_tools.py:44-57): oracle restatement and the CUDA kernel against golden
vectors of the reference's own `to_spectogram` (torchaudio MelSpectrogram + standardisation)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT
from oracle import lcasr_oracle as O

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_frontend import CASES, synth_wave  # noqa: E402  (the deterministic waveform generator only)


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: (z[k] if z[k].ndim else z[k].item()) for k in z.files}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_spectrogram_matches_reference_golden(name):
    g = _load(name)
    wav = synth_wave(int(g["channels"]), int(g["samples"]), int(g["seed"]))
    got = O.to_spectogram(wav.numpy(), bool(g["normalise"]))
    assert got.shape == g["spec"].shape
    assert np.abs(got - g["spec"]).max() < 2e-4 * np.abs(g["spec"]).max()


def test_host_tables_match_their_definitions():
    from lcasr_b200.frontend import build_tables
    cos_tab, sin_tab, fb = build_tables()
    assert cos_tab.shape == (512, 257) and fb.shape == (257, 80)
    assert float(fb.min()) >= 0 and int((fb.sum(0) > 0).sum()) == 80          # every mel bin has support
    assert float(cos_tab[:56].abs().max()) == 0 and float(cos_tab[456:].abs().max()) == 0  # window is centred, 400 wide
    x = torch.randn(512, dtype=torch.float64)
    win = torch.zeros(512, dtype=torch.float64)
    win[56:456] = torch.hann_window(400, periodic=True, dtype=torch.float64)
    ref = torch.fft.rfft(x * win)
    assert (x @ cos_tab - ref.real).abs().max() < 1e-9 and (x @ sin_tab - ref.imag).abs().max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_spectrogram_matches_reference_golden(cuda_device, name):
    from gpu_util import report
    from lcasr_b200.frontend import to_spectogram
    g = _load(name)
    wav = synth_wave(int(g["channels"]), int(g["samples"]), int(g["seed"])).to(cuda_device)
    got = to_spectogram(wav, global_normalisation=bool(g["normalise"])).cpu().numpy()
    assert got.shape == g["spec"].shape
    rel = float(np.abs(got - g["spec"]).max() / np.abs(g["spec"]).max())
    report(test="frontend", case=name, rel_max=rel)
    assert rel < 5e-4, f"mel spectrogram off by {rel} of the maximum"
    with pytest.raises(RuntimeError):
        to_spectogram(wav.cpu())
