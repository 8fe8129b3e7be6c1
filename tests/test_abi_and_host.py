"""CPU: the C-ABI library loads and exports every symbol include/lcasr_b200.h declares; host-side
logic of the drop-in classes (constructor kwargs, state_dict layout, error behaviour).  No kernels
are launched here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, GOLDEN_CASES, load_golden
from oracle import lcasr_oracle as O


def test_library_exports_every_declared_symbol():
    import lcasr_b200
    header = open(os.path.join(ROOT, "include", "lcasr_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(lcasr_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(lcasr_b200._lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert set(lcasr_b200._lib.EXPORTED_SYMBOLS) <= declared
    assert lib.lcasr_abi_version() == lcasr_b200._lib.ABI_VERSION == 3


def test_out_length_host_function():
    import lcasr_b200
    for T in (1, 2, 7, 8, 9, 264, 1000, 1024, 16384, 131072, 360000):
        assert lcasr_b200.ops.out_length(T) == O.calc_length(T)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_state_dict_layout_matches_reference(name):
    import lcasr_b200
    g = load_golden(name)
    cfg = O.make_config(**g["config"])
    model = lcasr_b200.SCConformerXL(**cfg)
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == g["state_dict_shapes"]
    sd = O.synth_state_dict(cfg, seed=g["weight_seed"], peak=g["peak"])
    model.load_state_dict(sd, strict=True)  # bin/load_pretrained.py:58 loads strict


def test_constructor_accepts_reference_kwargs_and_rejects_off_path_ones():
    import lcasr_b200
    # kwargs found in the released configs that the reference swallows through **kwargs
    m = lcasr_b200.SCConformerXL(vocab_size=127, n_layers=1, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                                 qk_rms_norm=False, shift_kvs=False, gated_sc=False, encoder_mode="conformer",
                                 self_condition_subsampling=False, ff_checkpoint_lvl=2, rotary_base_freq=1500000,
                                 flash_attn=True, checkpoint_every_n_layers=1, use_rotary=True, decoder_norm=True)
    assert m.decoder.num_classes == 128 and m.subsampling.subsampling_factor == 8
    assert float(m.rotary_pos_emb.rotary_interpolation_factor) == 1.0
    assert m.layers[0].attend.fn.n_heads == 2
    with pytest.raises(AssertionError):
        lcasr_b200.SCConformerXL(default_norm="batch_norm")
    with pytest.raises(NotImplementedError):
        lcasr_b200.SCConformerXL(subsampling="stacking")
    with pytest.raises(NotImplementedError):
        lcasr_b200.SCConformerXL(causal=True)
    # local attention is accepted (eval/run.py:42 sets config.model.attention_window_size), per-direction keys win
    w = lcasr_b200.SCConformerXL(n_layers=1, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                                 attention_window_size=128, attention_window_size_right=16)
    assert (w.layers[0].attend.fn.left_window, w.layers[0].attend.fn.right_window) == (128, 16)


def test_no_cpu_fallback():
    import lcasr_b200
    m = lcasr_b200.SCConformerXL(vocab_size=127, n_layers=1, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 80, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        lcasr_b200.GreedyCTCDecoder(None, 127)(torch.randn(5, 128))
    with pytest.raises(RuntimeError, match="CUDA"):
        lcasr_b200.CTCLoss(blank=127)(torch.randn(5, 1, 128), torch.zeros(1, 2, dtype=torch.long), torch.tensor([5]), torch.tensor([2]))


def test_eval_mode_with_gradients_is_opt_in_and_cuda_only():
    """dynamic_eval.py-style adaptation: model.grad_in_eval routes an eval()-mode call to the autograd path; like every other
    path it refuses CPU tensors instead of falling back"""
    import lcasr_b200
    m = lcasr_b200.SCConformerXL(vocab_size=127, n_layers=1, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32).eval()
    assert m.grad_in_eval is False
    m.grad_in_eval = True
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 80, 64), return_logits=True)
    with torch.no_grad():  # without gradients the fused inference call is taken (and refuses CPU tensors the same way)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(torch.randn(1, 80, 64))


def test_param_groups_follow_reference_rule():
    import lcasr_b200
    m = lcasr_b200.SCConformerXL(vocab_size=127, n_layers=1, d_model=64, n_heads=2, head_dim=32, subsampling_conv_channels=32,
                                 decoder_norm=True, use_rotary=True)
    groups = m.get_param_groups({"weight_decay": 0.1})
    n_all = sum(1 for _ in m.parameters())
    assert len(groups) == 2 and len(groups[0]["params"]) + len(groups[1]["params"]) == n_all
    assert groups[1]["weight_decay"] == 0.0
    assert not isinstance(m.get_param_groups({"weight_decay": 0.0}), list)


def test_bad_arguments_return_errors_not_crashes():
    import lcasr_b200
    L = lcasr_b200._lib
    assert L.lib.lcasr_gemm(None, None, 0, 4, 4, 4, None, 0, None, 1.0, None, 0, 0, None) == -1
    assert b"NULL" in L.lib.lcasr_last_error()
    assert L.lib.lcasr_layernorm(None, None, None, 1, 8, 1e-5, 0, None, None, 0, None) == -1
    assert L.lib.lcasr_model_workspace_bytes(None, 1, 100) == -1


def test_attention_tail_plan_host_logic():
    """the wave-quantisation planner of the dense attention launch is pure host code: a split is chosen only
    where the list-schedule model gains > 1.5 %, never for launches of at most one wave or short key ranges, and the
    chosen pieces always shorten the modelled makespan"""
    import ctypes as C
    import heapq
    from lcasr_b200 import _lib as L

    def plan(B, N, H, sms=148):
        t, p = C.c_int(-1), C.c_int(-1)
        L.call("lcasr_attention_tail_plan", B, N, H, sms, C.byref(t), C.byref(p))
        return t.value, p.value

    def makespan(W, full, pieces, piece_cost):
        h = [0.0] * W
        for _ in range(full):
            heapq.heapreplace(h, h[0] + 1.0)
        for _ in range(pieces):
            heapq.heapreplace(h, h[0] + piece_cost)
        return max(h)

    assert plan(1, 16384, 24) == (3, 2)      # BASELINE cfg 3: 1536 units = 10.38 waves
    assert plan(1, 45000, 16) == (1, 4)      # cfg 4: 2816 units = 19.03 waves
    assert plan(16, 2048, 6) == (5, 4)       # cfg 2
    assert plan(1, 128, 8) == (0, 0)         # cfg 1: fewer units than SMs
    assert plan(1, 1024, 24) == (0, 0)       # 8 key tiles: too short to cut
    assert plan(1, 37888, 1, sms=148) == (0, 0)  # exactly one full wave
    for B, N, H in ((1, 16384, 24), (1, 45000, 16), (16, 2048, 6), (2, 8192, 12), (3, 6000, 8)):
        t, P = plan(B, N, H)
        nq, nkt = -(-N // 256), -(-N // 128)
        U = B * H * nq
        if t:
            assert 1 <= t <= min(8, nq) and 2 <= P <= 4
            piece = (-(-nkt // P)) / nkt
            assert makespan(148, U - t * H, t * H * P, piece) < makespan(148, U, 0, 0.0)
    with pytest.raises(L.LcasrError):
        plan(0, 16384, 24)
    # random shapes: whatever is chosen stays inside the stated ranges and never lengthens the modelled launch
    import random
    rnd = random.Random(5)
    for _ in range(200):
        B, H = rnd.randint(1, 16), rnd.choice([1, 2, 6, 8, 12, 16, 24])
        N, sms = rnd.randint(1, 60000), rnd.choice([74, 132, 148])
        t, P = plan(B, N, H, sms)
        nq, nkt = -(-N // 256), -(-N // 128)
        U = B * H * nq
        if t == 0:
            assert P == 0
            continue
        assert U > sms and nkt >= 16 and 1 <= t <= min(8, nq) and 2 <= P <= 4 and nkt // P >= 4
        assert makespan(sms, U - t * H, t * H * P, (-(-nkt // P)) / nkt) < makespan(sms, U, 0, 0.0)


def test_workspace_size_host_functions():
    """sizes of the caller-owned workspaces are pure host arithmetic"""
    from lcasr_b200 import _lib as L
    f = L.lib.lcasr_attention_bwd_flash_workspace_bytes
    assert f(8, 2048, 6) == 8 * 6 * 2048 * 8          # float2 {lse2, D*scale} per (recording, head, token)
    assert f(1, 100, 2) == 1 * 2 * 128 * 8            # tokens padded to a multiple of 128
    assert f(0, 100, 2) == -1
    g = L.lib.lcasr_ctc_workspace_bytes
    assert g(1, 45000, 13500, 0) == 148 * 45000 * 16  # 1-hour lattice: one record per (chunk, frame), 148 chunks
    assert g(8, 2048, 614, 1) == 16 * 9 * 2048 * 16   # cfg 5, alpha and beta concurrently: 16 lattices x 9 chunks
    assert g(8, 100, 10, 1) == 0                      # 21 states: one chunk, the plain recursion is the same thing
