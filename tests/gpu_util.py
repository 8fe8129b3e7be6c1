"""Helpers shared by the -m gpu parity tests."""
import json
import os
import time

import torch

from conftest import ROOT, load_golden
from oracle import lcasr_oracle as O

REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def report(**kw):
    """Append a metric line to gpurun_out/parity_report.jsonl (scratch; summarised in profiles/)."""
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def build_model(g_or_cfg, device, compute_dtype="fp32", seed=12345, peak=1.0):
    import lcasr_b200
    if "config" in g_or_cfg and "weight_seed" in g_or_cfg:
        cfg = O.make_config(**g_or_cfg["config"])
        seed, peak = g_or_cfg["weight_seed"], g_or_cfg["peak"]
    else:
        cfg = g_or_cfg
    sd = O.synth_state_dict(cfg, seed=seed, peak=peak)
    model = lcasr_b200.SCConformerXL(**cfg, compute_dtype=compute_dtype)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval(), cfg, sd


def margin_mask(ref_lp: torch.Tensor, thresh: float) -> torch.Tensor:
    """frames whose top-1/top-2 margin in the fp32 reference exceeds `thresh`."""
    top2 = ref_lp.topk(2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) > thresh
