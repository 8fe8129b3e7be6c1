"""TEST INFRASTRUCTURE ONLY — golden vectors of the long-form moving-window path: the UNMODIFIED reference's
``lcasr.eval.utils.fetch_logits`` (utils.py:45-111) driving the unmodified reference model on CPU, fp32, followed by its
GreedyCTCDecoder (eval/run.py:84-89).  Also pins the oracle restatement (oracle.lcasr_oracle.fetch_logits).

Run in the build container:   python oracle/make_golden_longform.py"""
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
CFG = dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32, vocab_size=127)
# name -> (frames of the recording, seq_len, overlap)
CASES = {"longform_overlap875": (1500, 256, 224),     # 87.5 % overlap, ragged tail window
         "longform_overlap50_exact": (1024, 256, 128),  # windows tile the recording exactly
         "longform_single": (300, 512, 64)}             # seq_len > recording: one window, overlap forced to 0
# buffered mode (lcasr/eval/buffered_transcription.py): name -> (frames, seq_len (buffer), overlap (buffer - chunk))
BUFFERED_CASES = {"buffered_ragged": (1500, 256, 128),      # last chunk shorter, last buffer shifted inwards
                  "buffered_exact": (1024, 256, 64),        # chunks of 192 frames: 1024 is not a multiple -> ragged end
                  "buffered_single": (300, 512, 64)}        # buffer longer than the recording: one step, overlap 0


class _Tok:
    def __init__(self, v):
        self.v = v

    def vocab_size(self):
        return self.v


def main():
    SCConformerXL, GreedyCTCDecoder = load_reference()
    import contextlib
    import io
    from lcasr.eval.utils import fetch_logits  # the reference's own function
    torch.set_num_threads(8)
    cfg = O.make_config(**CFG)
    sd = O.synth_state_dict(cfg, seed=12345, peak=2.0)
    model = SCConformerXL(**cfg)
    model.load_state_dict(sd, strict=True)
    model.eval()
    model.device = "cpu"
    V = cfg["vocab_size"]
    for name, (frames, seq_len, overlap) in CASES.items():
        spec = O.synth_input(1, frames, cfg["feat_in"], seed=4321)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = fetch_logits(None, model, spec, seq_len, overlap, _Tok(V), use_tqdm=False)
        greedy = GreedyCTCDecoder(tokenizer=None, blank_id=V)(torch.as_tensor(ref))
        mine = O.fetch_logits(sd, cfg, spec, seq_len, overlap)
        err = np.abs(mine - ref).max()
        print(f"{name}: frames {frames} seq_len {seq_len} overlap {overlap} -> N={ref.shape[0]}; oracle-vs-reference max-abs {err:.2e}; "
              f"{len(greedy)} greedy tokens")
        assert mine.shape == ref.shape and err < 5e-5
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), config=json.dumps(CFG), frames=frames, seq_len=seq_len,
                            overlap=overlap, weight_seed=12345, peak=2.0, input_seed=4321, logits=ref.astype(np.float32),
                            greedy=np.array(greedy, dtype=np.int64))
    from lcasr.eval.buffered_transcription import fetch_logits as fetch_logits_buffered  # the reference's own function
    for name, (frames, seq_len, overlap) in BUFFERED_CASES.items():
        spec = O.synth_input(1, frames, cfg["feat_in"], seed=4321)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = fetch_logits_buffered(None, model, spec, seq_len, overlap, _Tok(V), use_tqdm=False)
        greedy = GreedyCTCDecoder(tokenizer=None, blank_id=V)(torch.as_tensor(ref))
        mine = O.fetch_logits_buffered(sd, cfg, spec, seq_len, overlap)
        err = np.abs(mine - ref).max()
        print(f"{name}: frames {frames} buffer {seq_len} overlap {overlap} -> N={ref.shape[0]}; oracle-vs-reference max-abs {err:.2e}; "
              f"{len(greedy)} greedy tokens")
        assert mine.shape == ref.shape and err < 5e-5
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), config=json.dumps(CFG), frames=frames, seq_len=seq_len,
                            overlap=overlap, weight_seed=12345, peak=2.0, input_seed=4321, logits=ref.astype(np.float32),
                            greedy=np.array(greedy, dtype=np.int64))


if __name__ == "__main__":
    main()
