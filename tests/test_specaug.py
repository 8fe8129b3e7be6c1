"""SpecAugment (lcasr/utils/augmentation.py:10-104): the oracle restatement, the consumption of the torch
random stream by the drop-in module, and the CUDA kernels, against golden vectors of the reference's own module
(oracle/make_golden_specaug.py records its output together with the uniform draws it consumed)."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import lcasr_oracle as O

CASES = ["specaug_iid_minp", "specaug_iid_zero", "specaug_shared", "specaug_no_time"]


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: (z[k] if z[k].ndim else z[k].item()) for k in z.files}
    g["kwargs"] = ast.literal_eval(str(g["kwargs"]))
    g["lengths"] = g["lengths"].tolist() or None
    g["spec"] = (O.synth_input(int(g["B"]), int(g["T"]), int(g["F"]), seed=int(g["input_seed"])) + 0.3)
    return g


@pytest.mark.parametrize("name", CASES)
def test_oracle_specaug_matches_reference_golden(name):
    g = _load(name)
    got = O.spec_augment(g["spec"].numpy(), g["lengths"], int(g["time_param"]), g["u_time"], int(g["freq_param"]), g["u_freq"],
                         zero_masking=g["kwargs"].get("zero_masking", False))
    assert np.abs(got - g["out"]).max() < 1e-6
    assert np.array_equal(got != g["spec"].numpy(), g["out"] != g["spec"].numpy())  # the same cells are masked


@pytest.mark.parametrize("name", CASES)
def test_module_consumes_the_random_stream_like_the_reference(name):
    """host logic: same effective mask parameters and — under the seed the golden run used — the same uniform draws in
    the same order, so a seeded training run masks the same cells as the reference"""
    from lcasr_b200.augmentation import SpecAugment
    g = _load(name)
    aug = SpecAugment(**g["kwargs"])
    assert aug.mask_params(int(g["F"]), int(g["T"])) == (int(g["time_param"]), int(g["freq_param"]))
    assert aug.mask_params(int(g["F"]), int(g["T"])) == O.specaug_params(
        int(g["T"]), int(g["F"]), g["kwargs"]["n_time_masks"], g["kwargs"]["n_freq_masks"], g["kwargs"]["freq_mask_param"],
        g["kwargs"].get("time_mask_param", -1), g["kwargs"].get("min_p", -1), g["kwargs"].get("max_p", 1.0))
    torch.manual_seed(int(g["torch_seed"]))
    tp, u_time, fp, u_freq = aug.draw(g["spec"])
    for mine, ref in ((u_time, g["u_time"]), (u_freq, g["u_freq"])):
        if ref.shape[0] == 0:
            assert mine is None
        else:
            assert np.array_equal(mine.numpy(), ref)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        aug(g["spec"])


def test_constructor_checks_follow_the_reference():
    from lcasr_b200.augmentation import SpecAugment
    with pytest.raises(AssertionError):
        SpecAugment(n_time_masks=2, n_freq_masks=0, freq_mask_param=0)          # neither min_p nor time_mask_param
    with pytest.raises(AssertionError):
        SpecAugment(n_time_masks=2, n_freq_masks=0, freq_mask_param=0, min_p=1.5)
    SpecAugment(n_time_masks=0, n_freq_masks=1, freq_mask_param=5, some_config_key=1)  # extra config keys are swallowed


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_specaug_kernel_matches_reference_golden(cuda_device, name):
    from lcasr_b200.augmentation import apply_masks
    g = _load(name)
    x = g["spec"].to(cuda_device)
    lens = None if g["lengths"] is None else torch.tensor(g["lengths"], device=cuda_device)
    dev = lambda u: None if u.shape[0] == 0 else torch.from_numpy(u).to(cuda_device)  # noqa: E731
    got = apply_masks(x, lens, int(g["time_param"]), dev(g["u_time"]), int(g["freq_param"]), dev(g["u_freq"]),
                      zero_masking=g["kwargs"].get("zero_masking", False)).cpu().numpy()
    ref, src = g["out"], g["spec"].numpy()
    masked = ref != src
    assert np.array_equal(got[~masked], src[~masked])          # untouched cells are bit-exact copies
    assert np.array_equal(got != src, masked)                  # the same cells are masked
    assert np.abs(got - ref).max() < 1e-6                      # fill value (mean of the un-padded frames) to fp32 rounding
    x_odd = x[:, :, :-3].contiguous()                          # T % 4 != 0: the scalar kernel gives the same cells
    if g["kwargs"]["n_time_masks"] == 0:
        got_odd = apply_masks(x_odd, None, int(g["time_param"]), None, int(g["freq_param"]), dev(g["u_freq"]), zero_masking=True)
        want = O.spec_augment(x_odd.cpu().numpy(), None, int(g["time_param"]), g["u_time"], int(g["freq_param"]), g["u_freq"], True)
        assert np.array_equal(got_odd.cpu().numpy(), want)


@pytest.mark.gpu
def test_specaug_module_on_cuda_matches_torchaudio_sequence(cuda_device):
    """the module draws from the CUDA generator exactly like the reference's chain of torchaudio calls: under the same
    seed the one-pass kernel output equals n_time + n_freq sequential torchaudio maskings"""
    import torchaudio.functional as AF
    from lcasr_b200.augmentation import SpecAugment
    B, F, T = 8, 80, 4096
    x = (O.synth_input(B, T, F, seed=5) + 0.25).to(cuda_device)
    lens = torch.tensor([4096, 4000, 3000, 4096, 2048, 4096, 1000, 4096], device=cuda_device)
    aug = SpecAugment(n_time_masks=10, n_freq_masks=2, freq_mask_param=27, min_p=0.05, max_p=1.0)
    torch.manual_seed(77)
    got = aug(x, lens)
    torch.manual_seed(77)
    valid = (torch.arange(T, device=cuda_device)[None, :] < lens[:, None])[:, None, :].expand(B, F, T)
    fill = x[valid].mean()
    tp, fp = aug.mask_params(F, T)
    y = x.unsqueeze(1)
    for _ in range(10):
        y = AF.mask_along_axis_iid(y, tp, fill, 3, p=1.0)
    for _ in range(2):
        y = AF.mask_along_axis_iid(y, fp, fill, 2, p=1.0)
    y = y.squeeze(1)
    assert torch.equal(got != x, y != x)
    assert (got - y).abs().max().item() < 1e-6
    assert 0.02 < float((got != x).float().mean()) < 0.9
