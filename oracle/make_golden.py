"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, CPU, fp32, eval) on deterministic synthetic weights/inputs.

Run in the build container:   python oracle/make_golden.py
The fixtures pin (a) the oracle restatement and (b) the CUDA path on the GPU box, where the reference
tree does not exist.  Each fixture stores the model kwargs, the seeds, the reference's
final_posteriors / length, the greedy token list produced by the reference's GreedyCTCDecoder, the
CTC loss from torch.nn.CTCLoss(blank=V, reduction='sum') exactly as exp/train.py:104,249 calls it,
and the reference's state_dict key->shape map.
"""
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

# name -> (model kwargs overrides, batch, frames, peak[, frames per recording of a ragged batch])
CASES = {
    "tiny_lengths": (dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                          vocab_size=127), 3, 400, 1.0, [400, 250, 333]),
    "tiny_lengths_dh128": (dict(n_layers=2, d_model=128, n_heads=1, head_dim=128, subsampling_conv_channels=64,
                                vocab_size=255), 2, 1500, 2.0, [1100, 1500]),
    "tiny_dh32_ragged": (dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                              vocab_size=127), 2, 264, 1.0),
    "tiny_dh128": (dict(n_layers=2, d_model=128, n_heads=1, head_dim=128, subsampling_conv_channels=64,
                        vocab_size=255), 1, 1000, 4.0),
    "tiny_rms_nosc": (dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                           vocab_size=127, default_norm="rms_norm", self_conditioning=False), 1, 400, 1.0),
    "tiny_norotary": (dict(n_layers=1, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                           vocab_size=127, use_rotary=False), 1, 333, 1.0),
    "cfg1_6L256D8H": (dict(O.BASELINE_MODELS["cfg1_6L256D8H"]), 1, 1024, 1.0),
    "cfg1_peaky": (dict(O.BASELINE_MODELS["cfg1_6L256D8H"]), 1, 1024, 12.0),
}


def run_case(name, overrides, batch, frames, peak, SCConformerXL, GreedyCTCDecoder, lengths=None):
    cfg = O.make_config(**overrides)
    sd = O.synth_state_dict(cfg, seed=12345, peak=peak)
    torch.manual_seed(12345)
    model = SCConformerXL(**cfg)
    ref_shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(sd, strict=True)  # bin/load_pretrained.py:58 loads strict
    model.eval()
    x = O.synth_input(batch, frames, cfg["feat_in"], seed=1234)
    len_t = None if lengths is None else torch.tensor(lengths, dtype=torch.long)
    with torch.no_grad():
        out = model(x) if lengths is None else model(x, length=len_t)  # eval/utils.py:84 calls positionally, length=None
        logits = model(x, length=len_t, return_logits=True)["final_posteriors"]
    lp, length = out["final_posteriors"], out["length"]
    V = cfg["vocab_size"]
    dec = GreedyCTCDecoder(tokenizer=None, blank_id=V)
    greedy = [dec(lp[b, : int(length[b])]) for b in range(batch)]
    N = lp.shape[1]
    tgt, tgt_len = O.synth_targets(batch, int(length.min()), vocab=V, frac=0.3, seed=99)
    ctc = torch.nn.CTCLoss(blank=V, reduction="sum")
    lp_g = lp.clone().requires_grad_(True)
    loss = ctc(lp_g.transpose(0, 1), tgt, length, tgt_len)
    loss.backward()
    per_sample = torch.nn.CTCLoss(blank=V, reduction="none")(lp.transpose(0, 1), tgt, length, tgt_len)

    # pin the oracle restatement while we are here
    o_lp, o_len = O.encoder_forward(sd, cfg, x, lengths=lengths)
    valid = (torch.arange(N)[None, :] < length[:, None]).unsqueeze(-1)  # padded rows are not part of the contract
    err = ((o_lp - lp) * valid).abs().max().item()
    o_greedy = [O.greedy_decode(lp[b, : int(length[b])], V) for b in range(batch)]
    o_ctc = O.ctc_loss(lp.numpy(), tgt.numpy(), length.numpy(), tgt_len.numpy(), V)
    print(f"{name}: N={N} oracle-vs-reference max-abs {err:.3e}; greedy equal {o_greedy == greedy}; "
          f"ctc ref {loss.item():.6f} oracle {o_ctc.sum():.6f}")
    assert err < 2e-5 * max(1.0, lp.abs().max().item() / 8) and o_greedy == greedy and torch.equal(o_len, length.to(torch.int32))
    assert abs(o_ctc.sum() - loss.item()) <= 1e-5 * abs(loss.item())

    grad = lp_g.grad
    g_idx = torch.Generator().manual_seed(7)
    flat_idx = torch.randint(0, grad.numel(), (4096,), generator=g_idx)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name + ".npz"),
        config=json.dumps(overrides), batch=batch, frames=frames, peak=peak,
        weight_seed=12345, input_seed=1234, target_seed=99, frame_lengths=np.array(lengths if lengths else [frames] * batch, dtype=np.int64),
        final_posteriors=lp.numpy().astype(np.float32),
        logits_sample=logits.numpy().astype(np.float32)[:, ::max(1, N // 8)],
        length=length.numpy().astype(np.int32),
        greedy=np.array([json.dumps(g) for g in greedy]),
        ctc_loss_sum=np.float64(loss.item()), ctc_loss_per_sample=per_sample.detach().numpy().astype(np.float64),
        ctc_grad_idx=flat_idx.numpy(), ctc_grad_sample=grad.reshape(-1)[flat_idx].numpy().astype(np.float32),
        ctc_grad_abs_sum=np.float64(grad.abs().sum().item()),
        state_dict_shapes=json.dumps(ref_shapes),
    )


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    SCConformerXL, GreedyCTCDecoder = load_reference()
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, (ov, b, t, peak, *rest) in CASES.items():
        if only and name not in only:
            continue
        run_case(name, ov, b, t, peak, SCConformerXL, GreedyCTCDecoder, lengths=rest[0] if rest else None)


if __name__ == "__main__":
    main()
