"""TEST INFRASTRUCTURE ONLY — golden vectors of the TRAINING step (cfg 5's path at test size) from the
UNMODIFIED reference (/root/reference, CPU, fp32, model.train()):

    out = model(audio_signal=x, length=None);  loss = CTCLoss(blank=V, reduction='sum')(...);  loss.backward()

exactly as exp/train.py:236-262 does it.  Stored per fixture: the loss, for EVERY parameter its gradient's
L2 norm and 256 sampled entries, the BatchRenorm running statistics after the step, and the train-mode
log-probs.  Also pins the oracle's restatement (oracle.lcasr_oracle.training_step) against the reference.

Run in the build container:   python oracle/make_golden_train.py
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import lcasr_oracle as O  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name -> (model kwargs, batch, frames, num_batches_tracked of every BatchRenorm before the step)
CASES = {
    "train_tiny_dh32": (dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                             vocab_size=127), 2, 264, 0),
    "train_tiny_dh128_nbt": (dict(n_layers=2, d_model=128, n_heads=1, head_dim=128, subsampling_conv_channels=64,
                                  vocab_size=255, decoder_norm=True), 2, 520, 20000),
    "train_rms_nosc_bias": (dict(n_layers=1, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                                 vocab_size=127, default_norm="rms_norm", self_conditioning=False, bias_in_ff=True,
                                 use_rotary=False), 3, 200, 9000),
    # padded batch (exp/train.py:236-241 passes length=a_lengths): pad masks in attention and the conv module
    "train_ragged_dh32": (dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                               vocab_size=127), 3, 264, 500, [264, 150, 201]),
    "train_ragged_dh128": (dict(n_layers=1, d_model=128, n_heads=1, head_dim=128, subsampling_conv_channels=64,
                                vocab_size=255, decoder_norm=True), 2, 1100, 0, [700, 1100]),
    # test-time adaptation (lcasr/eval/dynamic_eval.py:47-100, :217): model.eval() with gradients; the second one with
    # return_logits=True (the caller applies its own softmax)
    "train_evalmode_dh32": (dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                                 vocab_size=127), 3, 264, 700, None, dict(eval_mode=True)),
    "train_evalmode_logits_dh128": (dict(n_layers=1, d_model=128, n_heads=1, head_dim=128, subsampling_conv_channels=64,
                                         vocab_size=255, decoder_norm=True), 2, 520, 0, None, dict(eval_mode=True, return_logits=True)),
}


def run_case(name, overrides, batch, frames, nbt, SCConformerXL, lengths=None, opts=None):
    opts = opts or {}
    eval_mode, ret_logits = bool(opts.get("eval_mode")), bool(opts.get("return_logits"))
    import zlib
    cfg = O.make_config(**overrides)
    sd = O.synth_state_dict(cfg, seed=12345, peak=1.0)
    for k in sd:
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(nbt, dtype=torch.long)
    torch.manual_seed(12345)
    model = SCConformerXL(**cfg)
    model.load_state_dict(sd, strict=True)
    model.train(not eval_mode)
    if eval_mode:  # running statistics that differ from the initial (0, 1) so that they matter
        O.perturb_running_stats(sd, seed=5)
        model.load_state_dict(sd, strict=True)
    x = O.synth_input(batch, frames, cfg["feat_in"], seed=1234)
    V = cfg["vocab_size"]
    out = model(audio_signal=x, length=None if lengths is None else torch.tensor(lengths), return_logits=ret_logits)
    lp, length = out["final_posteriors"], out["length"]
    N = lp.shape[1]
    tgt, tgt_len = O.synth_targets(batch, N, vocab=V, frac=0.3, seed=99)
    if lengths is not None:  # transcripts no longer than 30 % of each recording's own token count
        tgt_len = torch.minimum(tgt_len, (0.3 * length.float()).long())
    lsm = torch.log_softmax(lp, dim=-1) if ret_logits else lp
    loss = torch.nn.CTCLoss(blank=V, reduction="sum")(lsm.transpose(0, 1), tgt, length, tgt_len).sum()
    loss.backward()
    ref_grads = {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None}  # unused parameters
    unused = [k for k, p in model.named_parameters() if p.grad is None]  # (e.g. decoder.reprojection without self-conditioning)
    ref_stats = {k: v.detach().clone() for k, v in model.state_dict().items() if k.endswith(("running_mean", "running_std"))}

    o_loss, o_grads, o_stats, o_lp = O.training_step(sd, cfg, x, tgt, tgt_len, lengths=lengths, train=not eval_mode,
                                                     return_logits=ret_logits)
    worst = 0.0
    floor = 1e-4 * max(g.norm().item() for g in ref_grads.values())  # mathematically-zero gradients (a bias in front of
    for k, g in ref_grads.items():                                  # a batch norm) are rounding noise on both sides
        rel = (o_grads[k] - g).norm().item() / max(g.norm().item(), floor)
        worst = max(worst, rel)
    stat_err = max(((o_stats[k] if k in o_stats else sd[k]) - ref_stats[k]).abs().max().item() for k in ref_stats)
    print(f"{name}: N={N} loss ref {loss.item():.6f} oracle {o_loss:.6f}; worst per-parameter grad rel-L2 {worst:.2e}; "
          f"running-stat max-abs {stat_err:.2e}; lp max-abs {(o_lp - lp.detach()).abs().max().item():.2e}")
    assert abs(o_loss - loss.item()) <= 1e-5 * abs(loss.item()) and worst < 2e-4 and stat_err < 1e-5

    store = dict(config=json.dumps(overrides), batch=batch, frames=frames, nbt=nbt, weight_seed=12345, input_seed=1234,
                 target_seed=99, loss=np.float64(loss.item()), frame_lengths=np.array(lengths if lengths else [], dtype=np.int64),
                 target_lengths=tgt_len.numpy().astype(np.int64), eval_mode=eval_mode, return_logits=ret_logits, log_probs=lp.detach().numpy().astype(np.float32),
                 length=length.numpy().astype(np.int32), names=np.array(list(ref_grads.keys())), unused=np.array(unused + [""]))
    for i, (k, g) in enumerate(ref_grads.items()):
        idx = torch.randint(0, g.numel(), (min(256, g.numel()),), generator=torch.Generator().manual_seed(zlib.crc32(k.encode()) & 0x7FFFFFFF))
        store[f"g{i}_norm"] = np.float64(g.norm().item())
        store[f"g{i}_idx"] = idx.numpy()
        store[f"g{i}_val"] = g.reshape(-1)[idx].numpy().astype(np.float32)
    for i, (k, v) in enumerate(ref_stats.items()):
        store[f"stat{i}"] = v.numpy().astype(np.float32)
    store["stat_names"] = np.array(list(ref_stats.keys()))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **store)


def main():
    SCConformerXL, _ = load_reference()
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, (ov, b, t, nbt, *rest) in CASES.items():
        if only and name not in only:
            continue
        run_case(name, ov, b, t, nbt, SCConformerXL, *rest)


if __name__ == "__main__":
    main()
