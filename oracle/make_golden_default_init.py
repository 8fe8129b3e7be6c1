"""TEST INFRASTRUCTURE ONLY — fixtures from the reference constructor's OWN default initialisation.

    python oracle/make_golden_default_init.py          (build container: needs the reference tree)

exp/train.py:363 seeds torch (`torch.manual_seed(12345)`) and builds the model with PyTorch's default
initialisers; SURVEY §8d prescribes exactly that for the parity weights.  The weights themselves are not
stored: `lcasr_b200.SCConformerXL(**cfg)` consumes the torch random stream in the same order as the reference
constructor (same sub-module construction order), so the same seed reproduces the same state_dict bit for bit —
the fixture pins that with a SHA-256 over the fp32 bytes of every tensor in sorted key order, checked on CPU
(`tests/test_default_init.py`) and again on the GPU box before the comparison.

Each fixture stores the reference's fp32 eval output (`final_posteriors`), its greedy tokens, the CTC loss
through `torch.nn.CTCLoss(blank=V, reduction='sum')`, and — for the record — how far the REFERENCE'S OWN bf16
path (`torch.autocast(bfloat16)`, what exp/train.py:244 runs) is from its fp32 output on the same input.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import lcasr_oracle as O  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
INIT_SEED = 12345  # exp/train.py:363

CASES = {
    # name -> (model kwargs overrides, batch, frames)
    "default_init_cfg1": (dict(O.BASELINE_MODELS["cfg1_6L256D8H"]), 1, 1024),
    "default_init_768d_dh128": (dict(n_layers=3, d_model=768, n_heads=6, head_dim=128, subsampling_conv_channels=256), 1, 1032),
}


def state_dict_sha256(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().to(torch.float32).contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    SCConformerXL, GreedyCTCDecoder = load_reference()
    torch.set_num_threads(8)
    for name, (ov, B, T) in CASES.items():
        cfg = O.make_config(**ov)
        torch.manual_seed(INIT_SEED)
        model = SCConformerXL(**cfg).eval()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        x = O.synth_input(B, T, cfg["feat_in"], seed=1234)
        with torch.no_grad():
            out = model(x)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                lo = model(x)["final_posteriors"].float()
        lp, length = out["final_posteriors"], out["length"]
        V = cfg["vocab_size"]
        dec = GreedyCTCDecoder(tokenizer=None, blank_id=V)
        greedy = [dec(lp[b]) for b in range(B)]
        tgt, tl = O.synth_targets(B, lp.shape[1], vocab=V, frac=0.3, seed=99)
        loss = torch.nn.CTCLoss(blank=V, reduction="sum")(lp.transpose(0, 1), tgt, length, tl).item()
        o_lp, _ = O.encoder_forward(sd, cfg, x)
        o_err = (o_lp - lp).abs().max().item()
        ref_bf16 = (lo - lp).abs().max().item()
        print(f"{name}: N={lp.shape[1]} |lp|max {lp.abs().max():.2f}; oracle-vs-reference {o_err:.2e}; reference bf16-autocast vs its fp32 "
              f"{ref_bf16:.3e}; ctc {loss:.4f}; greedy tokens {sum(len(g) for g in greedy)}")
        assert o_err < 2e-5 * max(1.0, lp.abs().max().item() / 8)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), config=json.dumps(ov), batch=B, frames=T,
                            init_seed=INIT_SEED, input_seed=1234, target_seed=99, weights_sha256=state_dict_sha256(sd),
                            final_posteriors=lp.numpy().astype(np.float32), length=length.numpy().astype(np.int32),
                            greedy=np.array([json.dumps(g) for g in greedy]), ctc_loss_sum=np.float64(loss),
                            ref_bf16_autocast_max_abs=np.float64(ref_bf16))


if __name__ == "__main__":
    main()
