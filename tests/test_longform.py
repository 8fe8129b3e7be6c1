"""Long-form moving-window inference (lcasr/eval/utils.py:45-111): CPU tests pin the oracle restatement
and the host-side window plan against golden vectors produced by the reference's own fetch_logits
(oracle/make_golden_longform.py); the GPU tests run the batched device path against the same vectors."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import lcasr_oracle as O

CASES = ["longform_overlap875", "longform_overlap50_exact", "longform_single"]


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["config"] = json.loads(str(g["config"]))
    return g


def _setup(g):
    cfg = O.make_config(**g["config"])
    sd = O.synth_state_dict(cfg, seed=int(g["weight_seed"]), peak=float(g["peak"]))
    spec = O.synth_input(1, int(g["frames"]), cfg["feat_in"], seed=int(g["input_seed"]))
    return cfg, sd, spec


@pytest.mark.parametrize("name", CASES)
def test_oracle_fetch_logits_matches_reference_golden(name):
    g = _load(name)
    cfg, sd, spec = _setup(g)
    got = O.fetch_logits(sd, cfg, spec, int(g["seq_len"]), int(g["overlap"]))
    assert got.shape == g["logits"].shape
    assert np.abs(got - g["logits"]).max() < 5e-5
    assert O.greedy_decode(torch.as_tensor(got), cfg["vocab_size"]) == g["greedy"].tolist()


@pytest.mark.parametrize("name", CASES)
def test_window_plan_matches_the_reference_loop(name):
    """the host-side plan (starts, lengths, merged positions, total frames) reproduces the reference's running
    `logit_position` bookkeeping, including the stop after the first short window"""
    from lcasr_b200.longform import plan_windows, window_positions
    g = _load(name)
    frames, seq_len, overlap = int(g["frames"]), int(g["seq_len"]), int(g["overlap"])
    wins, sl, ov = plan_windows(frames, seq_len, overlap, 8)
    ds = [O.calc_length(u) for _, u in wins]
    pos = window_positions(wins, ds, ov)
    assert max(p + n for p, n in zip(pos, ds)) == g["logits"].shape[0]
    assert all(b >= a for a, b in zip(pos, pos[1:]))
    lens = [u for _, u in wins]
    assert all(u == sl for u in lens[:-1]) and lens[-1] <= sl  # at most one short window, and it is the last one
    with pytest.raises(AssertionError):
        plan_windows(1000, 256, 100, 8)  # overlap not a multiple of the downsampling factor (utils.py:60)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_longform_device_path_matches_reference_golden(cuda_device, name, mode):
    import lcasr_b200
    from gpu_util import margin_mask, report
    from lcasr_b200.longform import fetch_logits, transcribe_longform
    g = _load(name)
    cfg, sd, spec = _setup(g)
    model = lcasr_b200.SCConformerXL(**cfg, compute_dtype=mode)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_device).eval()
    model.device = cuda_device
    ref = torch.from_numpy(g["logits"])
    for max_batch in (16, 3):  # grouping of the windows into forward batches must not matter
        got = torch.from_numpy(fetch_logits(None, model, spec, int(g["seq_len"]), int(g["overlap"]), None, use_tqdm=False,
                                            max_batch=max_batch))
        assert got.shape == ref.shape
        scale = max(1.0, ref.abs().max().item() / 8)
        err = (got - ref).abs().max().item()
        report(test="longform_" + mode, case=name, max_abs=err, max_batch=max_batch)
        assert err < (1e-4 if mode == "fp32" else 5e-2) * scale, f"{name}/{mode}: merged log-probs off by {err}"
    toks = transcribe_longform(model, spec.to(cuda_device), int(g["seq_len"]), int(g["overlap"]))
    if mode == "fp32":
        assert toks == g["greedy"].tolist()
    else:
        safe = margin_mask(ref, 4e-2 * scale)
        assert bool((got.argmax(-1) == ref.argmax(-1))[safe].all())


# ------------------------------------------------------------------------------------------------
# buffered mode (lcasr/eval/buffered_transcription.py:11-97); golden vectors from the reference's own function
# ------------------------------------------------------------------------------------------------
BUFFERED = ["buffered_ragged", "buffered_exact", "buffered_single"]


@pytest.mark.parametrize("name", BUFFERED)
def test_oracle_buffered_fetch_logits_matches_reference_golden(name):
    g = _load(name)
    cfg, sd, spec = _setup(g)
    got = O.fetch_logits_buffered(sd, cfg, spec, int(g["seq_len"]), int(g["overlap"]))
    assert got.shape == g["logits"].shape
    assert np.abs(got - g["logits"]).max() < 5e-5
    assert O.greedy_decode(torch.as_tensor(got), cfg["vocab_size"]) == g["greedy"].tolist()


def test_buffer_plan_host_logic():
    from lcasr_b200.longform import plan_buffers
    for frames, seq_len, overlap in [(1500, 256, 128), (1024, 256, 64), (300, 512, 64), (1024, 256, 0), (257, 256, 128),
                                     (4096, 1024, 512), (100, 100, 8), (2048, 512, 448)]:
        steps, sl, ov = plan_buffers(frames, seq_len, overlap)
        assert (steps, sl, ov) == O.buffered_positions(frames, seq_len, overlap)
        assert all(s1 - s0 == sl and 0 <= s0 and s1 <= frames for s0, s1, _, _ in steps)      # equal buffers, inside the recording
        assert all(s0 <= c0 and c1 <= s1 for s0, s1, c0, c1 in steps)                        # the chunk lies inside its buffer
        assert steps[0][2] == 0 and steps[-1][3] == frames                                   # the chunks tile the recording
        assert all(a[3] == b[2] for a, b in zip(steps, steps[1:]))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", BUFFERED)
def test_buffered_device_path_matches_reference_golden(cuda_device, name, mode):
    import lcasr_b200
    from gpu_util import margin_mask, report
    from lcasr_b200.longform import fetch_logits_buffered, transcribe_buffered
    g = _load(name)
    cfg, sd, spec = _setup(g)
    model = lcasr_b200.SCConformerXL(**cfg, compute_dtype=mode)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_device).eval()
    model.device = cuda_device
    ref = torch.from_numpy(g["logits"])
    outs = []
    for max_batch in (16, 2):
        got = torch.from_numpy(fetch_logits_buffered(None, model, spec, int(g["seq_len"]), int(g["overlap"]), None, use_tqdm=False,
                                                     max_batch=max_batch))
        assert got.shape == ref.shape
        scale = max(1.0, ref.abs().max().item() / 8)
        err = (got - ref).abs().max().item()
        report(test="buffered_" + mode, case=name, max_abs=err, max_batch=max_batch)
        assert err < (1e-4 if mode == "fp32" else 5e-2) * scale, f"{name}/{mode}: buffered log-probs off by {err}"
        outs.append(got)
    # the kept rows are copies of what a direct forward of the first buffer returns (bit-exact gather)
    first = model(spec[:, :, : min(int(g["seq_len"]), spec.shape[-1])].to(cuda_device))["final_posteriors"][0].cpu()
    n0 = min(first.shape[0], outs[0].shape[0], (int(g["seq_len"]) - int(g["overlap"])) // 8)
    assert torch.equal(outs[0][:n0], first[:n0])
    toks = transcribe_buffered(model, spec.to(cuda_device), int(g["seq_len"]), int(g["overlap"]))
    if mode == "fp32":
        assert toks == g["greedy"].tolist()
    else:
        safe = margin_mask(ref, 4e-2 * scale)
        assert bool((outs[0].argmax(-1) == ref.argmax(-1))[safe].all())
