"""TEST INFRASTRUCTURE ONLY — golden vectors for local (windowed) attention.  The reference's windowed path is flash-attn
CUDA only (attention.py:537 asserts it off on CPU), but the reference ships flash-attn's own pure-torch reference,
``attention_ref`` / ``construct_local_mask`` (lcasr/components/attention.py:330-420), which defines the window_size
semantics.  This script runs THAT function on CPU for a few shapes and stores q/k/v seeds + outputs, and pins the oracle's
band mask (oracle.lcasr_oracle.attention_forward with attention_window_size*) against it.
    python oracle/make_golden_window.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import lcasr_oracle as O  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
# name -> (B, N, H, Dh, (left, right))
CASES = {"window_dh32": (2, 300, 2, 32, (40, 40)), "window_dh128_asym": (1, 520, 1, 128, (64, 16)),
         "window_dh64_right_only": (1, 777, 2, 64, (-1, 100)), "window_dh32_wide": (1, 1000, 2, 32, (300, 300))}


def qkv(B, N, H, Dh, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, N, H, Dh, generator=g) for _ in range(3)]


def main():
    load_reference()
    from lcasr.components.attention import attention_ref
    for name, (B, N, H, Dh, win) in CASES.items():
        q, k, v = qkv(B, N, H, Dh, seed=N)
        out, _ = attention_ref(q, k, v, window_size=win, upcast=True)
        # the oracle's band mask through SDPA
        i, j = torch.arange(N)[:, None], torch.arange(N)[None, :]
        band = torch.ones(N, N, dtype=torch.bool)
        if win[0] >= 0:
            band &= j >= i - win[0]
        if win[1] >= 0:
            band &= j <= i + win[1]
        mask = torch.zeros(N, N).masked_fill(~band, float("-inf"))[None, None]
        mine = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                                attn_mask=mask).transpose(1, 2)
        err = (mine - out).abs().max().item()
        print(f"{name}: B={B} N={N} H={H} Dh={Dh} window={win}: band-mask SDPA vs reference attention_ref max-abs {err:.2e}")
        assert err < 2e-6
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), B=B, N=N, H=H, Dh=Dh, left=win[0], right=win[1], seed=N,
                            out=out.numpy().astype(np.float32))


if __name__ == "__main__":
    main()
