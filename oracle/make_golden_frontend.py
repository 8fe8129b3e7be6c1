"""TEST INFRASTRUCTURE ONLY — golden vectors of the front-end: the UNMODIFIED reference `to_spectogram`
(lcasr/utils/audio_tools.py:44-57, torchaudio MelSpectrogram + per-bin standardisation) on CPU.
    python oracle/make_golden_frontend.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import lcasr_oracle as O  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
# name -> (channels, samples, normalise)
CASES = {"frontend_2p5s": (1, 40000, True), "frontend_2ch_odd": (2, 116777, True), "frontend_raw_short": (1, 1700, False)}


def synth_wave(c, n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 16000.0
    tone = 0.2 * torch.sin(2 * np.pi * 440.0 * t) + 0.1 * torch.sin(2 * np.pi * 3100.0 * t * (1 + 0.1 * t))
    return (tone[None] + 0.05 * torch.randn(c, n, generator=g)).float()


def main():
    load_reference()
    from lcasr.utils.audio_tools import to_spectogram
    for name, (c, n, norm) in CASES.items():
        wav = synth_wave(c, n, seed=n)
        ref = to_spectogram(wav, global_normalisation=norm)
        mine = O.to_spectogram(wav.numpy(), global_normalisation=norm)
        scale = float(ref.abs().max())
        err = float(np.abs(mine - ref.numpy()).max()) / scale
        print(f"{name}: {tuple(ref.shape)}; oracle-vs-reference max-abs / max = {err:.2e}")
        assert mine.shape == tuple(ref.shape) and err < 2e-4
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), channels=c, samples=n, normalise=norm, seed=n,
                            spec=ref.numpy().astype(np.float32))


if __name__ == "__main__":
    main()
