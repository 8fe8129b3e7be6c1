"""The drop-in constructor reproduces the reference's default initialisation (exp/train.py:363 seeds torch with 12345 and
builds the model): same sub-module construction order => same consumption of the torch random stream => the same
state_dict bit for bit.  The fixtures (oracle/make_golden_default_init.py, generated from the UNMODIFIED reference) store a
SHA-256 of the reference's freshly constructed state_dict and its fp32 output; no weights are stored."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import lcasr_oracle as O

DEFAULT_INIT_CASES = ["default_init_cfg1", "default_init_768d_dh128"]


def state_dict_sha256(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().to(torch.float32).contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def build_default_init(g, compute_dtype="bf16"):
    import lcasr_b200
    cfg = O.make_config(**g["config"])
    torch.manual_seed(int(g["init_seed"]))
    model = lcasr_b200.SCConformerXL(**cfg, compute_dtype=compute_dtype)
    return model, cfg


@pytest.mark.parametrize("name", DEFAULT_INIT_CASES)
def test_constructor_reproduces_reference_default_init(name):
    g = load_golden(name)
    model, cfg = build_default_init(g)
    assert state_dict_sha256(model.state_dict()) == str(g["weights_sha256"])


@pytest.mark.parametrize("name", DEFAULT_INIT_CASES)
def test_oracle_matches_reference_on_default_init(name):
    g = load_golden(name)
    model, cfg = build_default_init(g)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    x = O.synth_input(g["batch"], g["frames"], cfg["feat_in"], seed=g["input_seed"])
    lp, length = O.encoder_forward(sd, cfg, x)
    ref = torch.from_numpy(g["final_posteriors"])
    assert (lp - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item() / 8)
    assert [O.greedy_decode(lp[b], cfg["vocab_size"]) for b in range(g["batch"])] == g["greedy"]
    tgt, tl = O.synth_targets(g["batch"], lp.shape[1], vocab=cfg["vocab_size"], frac=0.3, seed=g["target_seed"])
    nll = O.ctc_loss(lp.numpy(), tgt.numpy(), length.numpy(), tl.numpy(), cfg["vocab_size"]).sum()
    assert abs(nll - float(g["ctc_loss_sum"])) < 1e-5 * abs(float(g["ctc_loss_sum"]))
    # for the record: the reference's OWN bf16 path (torch.autocast) is further from its fp32 output than north_star's 2e-2
    assert float(g["ref_bf16_autocast_max_abs"]) > 2e-2
