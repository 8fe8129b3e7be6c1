"""Optimizer step: oracle restatement and the device multi-tensor kernels against golden vectors of the
unmodified reference MADGRAD + clip_grad_norm_ (oracle/make_golden_optim.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import lcasr_oracle as O

CASES = ["optim_momentum", "optim_nomomentum_wd", "optim_decoupled_two_groups"]


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["config"], g["shapes"] = json.loads(str(g["config"])), [tuple(s) for s in json.loads(str(g["shapes"]))]
    return g


def synth(step, i, shape, seed=0):  # same generator as the golden script
    g = torch.Generator().manual_seed(1000 * seed + 10 * step + i)
    return torch.randn(shape, generator=g) * (0.3 + 0.2 * i)


@pytest.mark.parametrize("name", CASES)
def test_oracle_madgrad_matches_reference_golden(name):
    g = _load(name)
    c, shapes = g["config"], g["shapes"]
    ps = [synth(-1, i, s, seed=7).numpy() for i, s in enumerate(shapes)]
    st = [dict() for _ in shapes]
    for step in range(c["steps"]):
        grads = [None if (step == 1 and i == 2) else synth(step, i, s).numpy() for i, s in enumerate(shapes)]
        live = [i for i, x in enumerate(grads) if x is not None]
        gs = [grads[i] for i in live]
        if c["clip"] > 0:
            gs, _ = O.clip_grad_norm(gs, c["clip"])
        for i, gg in zip(live, gs):
            wd = 0.0 if ("two_groups" in name and i >= 3) else c["weight_decay"]
            ps[i] = O.madgrad_step(ps[i], gg, st[i], step, c["lr"], c["momentum"], wd, c["eps"], c["decouple_decay"])
    for i in range(len(shapes)):
        assert np.abs(ps[i] - g[f"p{i}"]).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("fused_clip", [True, False])
@pytest.mark.parametrize("name", CASES)
def test_device_madgrad_matches_reference_golden(cuda_device, name, fused_clip):
    from lcasr_b200 import optim
    g = _load(name)
    c, shapes = g["config"], g["shapes"]
    params = [torch.nn.Parameter(synth(-1, i, s, seed=7).to(cuda_device)) for i, s in enumerate(shapes)]
    kw = dict(lr=c["lr"], momentum=c["momentum"], weight_decay=c["weight_decay"], eps=c["eps"], decouple_decay=c["decouple_decay"])
    if "two_groups" in name:
        opt = optim.MADGRAD([{"params": params[:3]}, {"params": params[3:], "weight_decay": 0.0}], **kw)
    else:
        opt = optim.MADGRAD(params, **kw)
    if fused_clip:
        opt.max_grad_norm = c["clip"]
    for step in range(c["steps"]):
        for i, (p, s) in enumerate(zip(params, shapes)):
            p.grad = None if (step == 1 and i == 2) else synth(step, i, s).to(cuda_device)
        if not fused_clip and c["clip"] > 0:
            total = optim.clip_grad_norm_(params, c["clip"])
            ref_total = torch.sqrt(sum((synth(step, i, s) ** 2).sum() for i, s in enumerate(shapes) if not (step == 1 and i == 2)))
            assert abs(total.item() - ref_total.item()) < 1e-4 * ref_total.item()
        opt.step()
        opt.zero_grad()
    for i, p in enumerate(params):
        assert (p.detach().cpu() - torch.from_numpy(g[f"p{i}"])).abs().max().item() < 2e-6, f"parameter {i}"
        assert (opt.state[p]["grad_sum_sq"].cpu() - torch.from_numpy(g[f"gss{i}"])).abs().max().item() < 1e-5 * max(1.0, float(np.abs(g[f"gss{i}"]).max()))
        assert (opt.state[p]["s"].cpu() - torch.from_numpy(g[f"s{i}"])).abs().max().item() < 1e-5 * max(1.0, float(np.abs(g[f"s{i}"]).max()))
    assert int(opt.state["k"]) == c["steps"]


@pytest.mark.gpu
def test_training_loop_with_device_optimizer(cuda_device):
    """exp/train.py's step end to end on the drop-ins: forward, CTC loss, backward, clip + MADGRAD on the device"""
    import lcasr_b200
    cfg = O.make_config(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32, vocab_size=127)
    model = lcasr_b200.SCConformerXL(**cfg)
    model.load_state_dict(O.synth_state_dict(cfg, seed=12345), strict=True)
    model = model.to(cuda_device).train()
    opt = lcasr_b200.optim.MADGRAD(model.get_param_groups({"weight_decay": 1e-4}), lr=3e-3)
    opt.max_grad_norm = 0.8
    ctc = lcasr_b200.CTCLoss(blank=127, reduction="sum")
    x = O.synth_input(2, 264, 80, seed=1).to(cuda_device)
    losses = []
    for _ in range(6):
        out = model(x)
        tgt, tl = O.synth_targets(2, out["final_posteriors"].shape[1], vocab=127)
        loss = ctc(out["final_posteriors"].transpose(0, 1), tgt, out["length"], tl)
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
