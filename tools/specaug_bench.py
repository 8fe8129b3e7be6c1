"""SpecAugment at cfg-5 batch shape (B=8, 80 bins, 16384 frames): the one-pass kernels vs the reference's chain of torchaudio
maskings on the same GPU (n_time + n_freq full torch.where passes).  CUDA events, median of 20."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchaudio.functional as AF
from lcasr_b200.augmentation import SpecAugment
dev = torch.device("cuda", 0)
B, F, T = 8, 80, 16384
x = torch.randn(B, F, T, device=dev)
lens = torch.full((B,), T, device=dev)
aug = SpecAugment(n_time_masks=10, n_freq_masks=2, freq_mask_param=27, min_p=0.05, max_p=1.0)
tp, fp = aug.mask_params(F, T)


def ref_chain():  # lcasr/utils/augmentation.py:73-97 on the GPU
    valid = (torch.arange(T, device=dev)[None, :] < lens[:, None])[:, None, :].expand(B, F, T)
    fill = x[valid].mean()
    y = x.unsqueeze(1)
    for _ in range(10):
        y = AF.mask_along_axis_iid(y, tp, fill, 3, p=1.0)
    for _ in range(2):
        y = AF.mask_along_axis_iid(y, fp, fill, 2, p=1.0)
    return y.squeeze(1)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2] * 1e3


mine, ref = timeit(lambda: aug(x, lens)), timeit(ref_chain)
print(json.dumps({"op": "specaugment", "shape": [B, F, T], "n_time_masks": 10, "n_freq_masks": 2, "lcasr_b200_us": round(mine, 1),
                  "torchaudio_chain_us": round(ref, 1), "algorithmic_MB": round(3 * B * F * T * 4 / 1e6, 1),
                  "note": "includes the device draws of the uniform numbers (24 torch.rand calls) in both arms"}))
