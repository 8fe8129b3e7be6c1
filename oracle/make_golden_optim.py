"""TEST INFRASTRUCTURE ONLY — golden vectors of the optimizer step: the UNMODIFIED reference MADGRAD
(lcasr/optim/madgrad.py) preceded by torch.nn.utils.clip_grad_norm_ exactly like exp/train.py:54-56, on CPU, fp32, for a
few steps of deterministic gradients.  Also pins the oracle restatement.   python oracle/make_golden_optim.py"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import lcasr_oracle as O  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
SHAPES = [(300,), (64, 130), (5,), (3, 7, 11), (4097,)]
CASES = {
    "optim_momentum": dict(lr=1e-2, momentum=0.9, weight_decay=0.0, eps=1e-6, decouple_decay=False, clip=0.8, steps=5),
    "optim_nomomentum_wd": dict(lr=3e-3, momentum=0.0, weight_decay=0.01, eps=1e-6, decouple_decay=False, clip=0.0, steps=4),
    "optim_decoupled_two_groups": dict(lr=1e-2, momentum=0.9, weight_decay=0.02, eps=1e-6, decouple_decay=True, clip=0.5, steps=3),
}


def synth(step, i, shape, seed=0):
    g = torch.Generator().manual_seed(1000 * seed + 10 * step + i)
    return torch.randn(shape, generator=g) * (0.3 + 0.2 * i)


def main():
    spec = importlib.util.spec_from_file_location("ref_madgrad", "/root/reference/lcasr/optim/madgrad.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name, c in CASES.items():
        params = [torch.nn.Parameter(synth(-1, i, s, seed=7)) for i, s in enumerate(SHAPES)]
        kw = dict(lr=c["lr"], momentum=c["momentum"], weight_decay=c["weight_decay"], eps=c["eps"], decouple_decay=c["decouple_decay"])
        if "two_groups" in name:  # a no-decay group like BaseModel.get_param_groups builds (base.py:25-68)
            opt = mod.MADGRAD([{"params": params[:3]}, {"params": params[3:], "weight_decay": 0.0}], **kw)
        else:
            opt = mod.MADGRAD(params, **kw)
        o_p = [p.detach().numpy().copy() for p in params]
        o_state = [dict() for _ in params]
        for step in range(c["steps"]):
            grads = [synth(step, i, s) for i, s in enumerate(SHAPES)]
            if step == 1:
                grads[2] = None  # a parameter without gradient is skipped
            for p, g in zip(params, grads):
                p.grad = None if g is None else g.clone()
            if c["clip"] > 0:
                torch.nn.utils.clip_grad_norm_(params, c["clip"])
            opt.step()
            # oracle
            live = [i for i, g in enumerate(grads) if g is not None]
            gs = [grads[i].numpy() for i in live]
            if c["clip"] > 0:
                gs, _ = O.clip_grad_norm(gs, c["clip"])
            for i, g in zip(live, gs):
                wd = 0.0 if ("two_groups" in name and i >= 3) else c["weight_decay"]
                o_p[i] = O.madgrad_step(o_p[i], g, o_state[i], step, c["lr"], c["momentum"], wd, c["eps"], c["decouple_decay"])
        err = max(np.abs(a - p.detach().numpy()).max() for a, p in zip(o_p, params))
        print(f"{name}: oracle-vs-reference max-abs on parameters after {c['steps']} steps: {err:.2e}")
        assert err < 1e-6
        store = dict(config=json.dumps(c), shapes=json.dumps(SHAPES))
        for i, p in enumerate(params):
            store[f"p{i}"] = p.detach().numpy()
            store[f"gss{i}"] = opt.state[p]["grad_sum_sq"].numpy()
            store[f"s{i}"] = opt.state[p]["s"].numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **store)


if __name__ == "__main__":
    main()
