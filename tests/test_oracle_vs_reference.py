"""CPU, build container only: the oracle against the UNMODIFIED reference imported from
/root/reference (skipped on machines without the reference tree, e.g. the GPU box)."""
import pytest
import torch

from oracle import lcasr_oracle as O
from oracle.ref_import import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("overrides,batch,frames", [
    (dict(n_layers=3, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32, vocab_size=127), 2, 520),
    (dict(n_layers=1, d_model=128, n_heads=1, head_dim=128, subsampling_conv_channels=32, vocab_size=63, bias_in_ff=True), 1, 300),
    (dict(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32, vocab_size=127, decoder_norm=False,
          legasee_double_norm=False), 1, 257),
])
def test_oracle_equals_reference(overrides, batch, frames):
    SCConformerXL, GreedyCTCDecoder = load_reference()
    cfg = O.make_config(**overrides)
    sd = O.synth_state_dict(cfg, seed=777)
    model = SCConformerXL(**cfg)
    model.load_state_dict(sd, strict=True)
    model.eval()
    x = O.synth_input(batch, frames, seed=5)
    with torch.no_grad():
        ref = model(x)
    lp, length = O.encoder_forward(sd, cfg, x)
    assert (lp - ref["final_posteriors"]).abs().max().item() < 2e-5
    assert length.tolist() == ref["length"].tolist()
    dec = GreedyCTCDecoder(None, blank_id=cfg["vocab_size"])
    for b in range(batch):
        assert dec(ref["final_posteriors"][b]) == O.greedy_decode(lp[b], cfg["vocab_size"])


def test_state_dict_layout_equals_reference():
    SCConformerXL, _ = load_reference()
    for overrides in (dict(), dict(default_norm="rms_norm"), dict(bias_in_ff=True), dict(use_rotary=False, decoder_norm=False)):
        cfg = O.make_config(n_layers=2, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32, vocab_size=127, **overrides)
        ref = {k: tuple(v.shape) for k, v in SCConformerXL(**cfg).state_dict().items()}
        assert ref == {k: tuple(v) for k, v in O.state_dict_shapes(cfg).items()}
