"""CPU: the oracle's TRAINING restatement (oracle.lcasr_oracle.training_step: train-mode / eval-mode-with-gradients forward,
CTC loss, autograd) against the golden vectors the unmodified reference produced (oracle/make_golden_train.py): the loss,
every parameter gradient (norm + 256 sampled entries) and the BatchRenorm buffers after the step."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import lcasr_oracle as O

CASES = ["train_tiny_dh32", "train_tiny_dh128_nbt", "train_rms_nosc_bias", "train_ragged_dh32", "train_ragged_dh128",
         "train_evalmode_dh32", "train_evalmode_logits_dh128"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_training_step_matches_reference_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    cfg = O.make_config(**json.loads(str(g["config"])))
    sd = O.synth_state_dict(cfg, seed=int(g["weight_seed"]), peak=1.0)
    for k in sd:
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(int(g["nbt"]), dtype=torch.long)
    eval_mode = bool(g["eval_mode"]) if "eval_mode" in g else False
    ret_logits = bool(g["return_logits"]) if "return_logits" in g else False
    if eval_mode:
        O.perturb_running_stats(sd, seed=5)
    x = O.synth_input(int(g["batch"]), int(g["frames"]), cfg["feat_in"], seed=int(g["input_seed"]))
    lengths = g["frame_lengths"].tolist() if "frame_lengths" in g and g["frame_lengths"].size else None
    N = O.calc_length(int(g["frames"]))
    tgt, tl = O.synth_targets(int(g["batch"]), N, vocab=cfg["vocab_size"], frac=0.3, seed=int(g["target_seed"]))
    if "target_lengths" in g:
        tl = torch.from_numpy(g["target_lengths"])
    loss, grads, stats, lp = O.training_step(sd, cfg, x, tgt, tl, lengths=lengths, train=not eval_mode, return_logits=ret_logits)
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert np.abs(lp.numpy() - g["log_probs"]).max() < 5e-5
    names = [str(n) for n in g["names"]]
    norms = [float(g[f"g{i}_norm"]) for i in range(len(names))]
    floor = 1e-4 * max(norms)  # mathematically-zero gradients (a bias in front of a batch norm) are rounding noise
    for i, n in enumerate(names):
        got = grads[n].reshape(-1)
        idx, val = torch.from_numpy(g[f"g{i}_idx"]), torch.from_numpy(g[f"g{i}_val"])
        assert abs(got.norm().item() - norms[i]) <= 2e-4 * max(norms[i], floor), n
        assert (got[idx] - val).norm().item() <= 2e-4 * max(val.norm().item(), floor), n
    for n in (str(u) for u in g["unused"] if str(u)):
        assert n not in grads, f"{n}: unused in the reference, no gradient expected"
    for i, n in enumerate(str(s) for s in g["stat_names"]):
        after = stats[n] if n in stats else sd[n]  # eval mode leaves the buffers alone
        assert (after - torch.from_numpy(g[f"stat{i}"])).abs().max().item() < 1e-5, n
