"""ctypes binding of liblcasr_b200.so (the C ABI declared in include/lcasr_b200.h).

There is deliberately no fallback: if the shared library is missing the import raises, and every
wrapper raises ``RuntimeError`` with ``lcasr_last_error()`` when a call returns non-zero.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LCASR_LIB_PATH") or os.path.join(_HERE, "liblcasr_b200.so")  # env: A/B runs of two builds

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU_TANH, ACT_SILU = 0, 1, 2
NORM_LAYERNORM, NORM_RMSNORM = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2
ATTN_AUTO, ATTN_SIMT, ATTN_TCGEN05 = 0, 1, 2
ABI_VERSION = 3

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class LcasrConfig(C.Structure):
    _fields_ = [(n, i32) for n in (
        "abi_version", "n_layers", "d_model", "n_heads", "head_dim", "feat_in", "conv_channels",
        "conv_kernel_size", "num_classes", "norm_kind", "decoder_norm", "use_rotary", "self_conditioning",
        "legasee_double_norm", "bias_in_ff", "compute_dtype")] + [("rotary_interp", f32), ("norm_eps", f32),
                                                                    ("attn_window_left", i32), ("attn_window_right", i32)]


LAYER_FIELDS = (
    "ff1_norm_w", "ff1_norm_b", "ff1_fc1_w", "ff1_fc1_b", "ff1_fc2_w", "ff1_fc2_b",
    "attn_norm_w", "attn_norm_b", "qkv_w", "out_w",
    "conv_norm_w", "conv_norm_b", "pw1_w", "pw1_b",
    "dw_w", "dw_b", "brn_mean", "brn_std", "brn_w", "brn_b", "pw2_w", "pw2_b",
    "ff2_norm_w", "ff2_norm_b", "ff2_fc1_w", "ff2_fc1_b", "ff2_fc2_w", "ff2_fc2_b",
    "norm_out_w", "norm_out_b", "qkv_w_il", "pw1_w_glu", "pw1_b_glu")


class LcasrLayerWeights(C.Structure):
    _fields_ = [(n, vp) for n in LAYER_FIELDS]


MODEL_FIELDS = (
    "conv0_w", "conv0_b", "dw1_w", "dw1_b", "pw1_w", "pw1_b", "dw2_w", "dw2_b", "pw2_w", "pw2_b",
    "sub_out_w", "inv_freq", "dec_norm_w", "dec_norm_b", "dec_ff_w", "dec_ff_b", "dec_rep_w", "dec_rep_b")


class LcasrWeights(C.Structure):
    _fields_ = [(n, vp) for n in MODEL_FIELDS] + [("layers_host", C.POINTER(LcasrLayerWeights))]


EPI_SCALE, EPI_EXP2, EPI_DS, EPI_GELU_BWD, EPI_SILU_BWD = 0, 1, 2, 3, 4


class LcasrGemmExArgs(C.Structure):
    """mirror of lcasr_gemm_ex_args (include/lcasr_b200.h)"""
    _fields_ = ([(n, vp) for n in ("A", "B", "out", "aux", "rowvec")] + [("M", i64), ("N", i32), ("K", i32),
                ("a_mn", i32), ("b_mn", i32)] + [(n, i64) for n in ("lda", "ldb", "ldo", "ldaux")] +
                [("nb1", i32), ("nb2", i32)] +
                [(n, i64) for n in ("sa1", "sa2", "sb1", "sb2", "so1", "so2", "sx1", "sx2", "sr1", "sr2")] +
                [("alpha", f32), ("epi", i32), ("out_dtype", i32), ("ksplit", i32)])


class LcasrOptTensor(C.Structure):
    """mirror of lcasr_opt_tensor"""
    _fields_ = [("p", vp), ("g", vp), ("gss", vp), ("s", vp), ("x0", vp), ("n", i64)]


# name -> argtypes; every function returns int status unless listed in _OTHER_RESTYPE
_SIGNATURES = {
    "lcasr_layernorm": [vp, vp, vp, i64, i32, f32, i32, vp, vp, i32, vp],
    "lcasr_layernorm_chain": [vp, i32, vp, vp, i64, i32, f32, i32, i32, vp, vp, i32, vp],
    "lcasr_subsample_conv0": [vp, vp, vp, i32, i32, i64, i32, vp, i32, vp],
    "lcasr_subsample_dwconv": [vp, i32, vp, vp, i32, i64, i32, i32, vp, vp],
    "lcasr_subsample_conv0_dw": [vp, vp, vp, vp, vp, i32, i32, i64, i32, vp, vp],
    "lcasr_gemm": [vp, vp, i32, i64, i32, i32, vp, i32, vp, f32, vp, i32, i32, vp],
    "lcasr_cast_f32": [vp, i64, vp, i32, vp],
    "lcasr_gemm_rope": [vp, vp, i64, i32, i32, vp, vp, i64, i32, i32, vp, vp],
    "lcasr_gemm_glu": [vp, vp, i64, i32, i32, vp, vp, vp],
    "lcasr_attention_qkv": [vp, i32, i64, vp, i32, i32, vp, vp],
    "lcasr_glu": [vp, i32, i64, i32, vp, vp],
    "lcasr_rope_table": [vp, f32, i64, i64, i32, vp, vp, vp],
    "lcasr_rope_table_t": [vp, f32, i64, i64, i32, vp, vp, vp],
    "lcasr_rope_split": [vp, i32, i32, i64, i32, i32, vp, vp, vp, vp, vp, i32, i64, vp],
    "lcasr_attention": [vp, vp, vp, i32, i32, i64, i32, i32, i32, i64, vp, i32, vp],
    "lcasr_attention_cross": [vp, vp, vp, i32, i32, i64, i64, i32, i32, vp, i32, vp],
    "lcasr_attention_masked": [vp, vp, vp, i32, i32, i64, i64, vp, i32, i32, vp, i32, vp],
    "lcasr_attention_window": [vp, vp, vp, i32, i32, i64, vp, i32, i32, i32, i32, vp, i32, vp],
    "lcasr_glu_masked": [vp, i32, i32, i64, i32, vp, vp, vp],
    "lcasr_dwconv_brn_silu": [vp, i32, i32, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "lcasr_softmax": [vp, i32, i64, i32, vp, i32, vp],
    "lcasr_log_softmax_argmax": [vp, i64, i32, vp, vp],
    "lcasr_argmax_rows": [vp, i64, i32, vp, vp],
    "lcasr_greedy_collapse": [vp, i32, i64, vp, i32, vp, vp, vp],
    "lcasr_ctc_loss_fwd": [vp, i32, i64, i32, vp, i64, vp, vp, i32, vp, vp, vp],
    "lcasr_ctc_loss_bwd": [vp, i32, i64, i32, vp, i64, vp, vp, i32, vp, vp, vp, vp, vp, vp],
    "lcasr_grad_sumsq": [vp, vp, vp, i32, vp, vp],
    "lcasr_grad_scale": [vp, vp, vp, i32, vp, f32, vp],
    "lcasr_madgrad_step": [vp, vp, vp, i32, vp, f32, f32, f32, f32, f32, f32, i32, vp],
    "lcasr_melspec": [vp, i32, i64, vp, vp, vp, i32, vp, vp, i32, vp],
    "lcasr_window_merge": [vp, i32, i32, vp, vp, vp, i32, i64, vp, vp, vp],
    "lcasr_attention_train_masked": [vp, vp, vp, i32, i64, vp, i32, i32, vp, vp, vp],
    "lcasr_subsample_l1_bwd": [vp, vp, vp, vp, vp, i32, i32, i64, i32, vp, vp, vp, vp, vp],
    "lcasr_mask_rows": [vp, i32, i32, i64, i32, vp, vp],
    "lcasr_window_concat": [vp, i32, i32, vp, vp, vp, i32, i64, vp, vp, vp],
    "lcasr_specaug_mean": [vp, i32, i32, i64, vp, vp, vp],
    "lcasr_specaug_apply": [vp, i32, i32, i64, i32, i32, vp, i32, i32, vp, i32, vp, vp, vp],
    "lcasr_ctc_loss_fwd_ab": [vp, i32, i64, i32, vp, i64, vp, vp, i32, vp, vp, vp, vp],
    "lcasr_ctc_loss_grad": [vp, i32, i64, i32, vp, i64, vp, vp, i32, vp, vp, vp, vp, vp, vp],
    "lcasr_gemm_ex": [C.POINTER(LcasrGemmExArgs), vp],
    "lcasr_attention_bwd_pds": [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, vp, vp, vp],
    "lcasr_attention_bwd_flash": [vp, vp, vp, vp, vp, vp, i32, i64, i64, i64, i32, i32, vp, vp, vp, vp, i64, vp],
    "lcasr_attention_train": [vp, vp, vp, i32, i64, i32, i32, vp, vp, vp],
    "lcasr_gemm_act_pre": [vp, vp, i64, i32, i32, vp, i32, vp, vp, vp],
    "lcasr_scale_cast": [vp, i64, f32, vp, vp],
    "lcasr_act_fwd": [vp, i64, i32, vp, vp],
    "lcasr_act_bwd": [vp, vp, i64, i32, vp, vp],
    "lcasr_add_bf16": [vp, vp, i64, vp],
    "lcasr_glu_bwd": [vp, vp, i64, i32, vp, vp],
    "lcasr_rope_bwd_merge": [vp, vp, vp, i32, i64, i32, i32, vp, vp, vp, vp],
    "lcasr_rowdot": [vp, vp, i32, i64, i32, i32, vp, vp],
    "lcasr_softmax_bwd": [vp, vp, i64, i32, f32, vp, vp],
    "lcasr_log_softmax_bwd": [vp, vp, i64, i32, f32, vp, vp],
    "lcasr_colsum": [vp, i32, i64, i32, f32, vp, vp],
    "lcasr_layernorm_bwd": [vp, vp, i32, vp, i64, i32, f32, i32, i32, vp, vp, vp, vp],
    "lcasr_layernorm_bwd_cast": [vp, vp, i32, vp, i64, i32, f32, i32, i32, vp, vp, vp, vp, f32, vp],
    "lcasr_dwconv1d_fwd": [vp, i32, i64, i32, i32, vp, vp, vp, vp, vp, vp],
    "lcasr_dwconv1d_bwd_data": [vp, i32, i64, i32, i32, vp, vp, vp],
    "lcasr_dwconv1d_bwd_weight": [vp, vp, i32, i64, i32, i32, vp, vp, vp],
    "lcasr_brn_train_stats": [vp, vp, i64, i32, vp, vp, f32, f32, f32, f32, vp, vp, vp, vp, vp, vp],
    "lcasr_affine_silu": [vp, i64, i32, vp, vp, vp, vp],
    "lcasr_affine_silu_bwd": [vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp],
    "lcasr_brn_bwd_finalize": [vp, vp, i64, i32, vp, vp, vp, vp, vp, vp],
    "lcasr_affine3": [vp, vp, i64, i32, vp, vp, vp],
    "lcasr_subsample_dwconv_bwd_data": [vp, vp, i32, i64, i32, i32, vp, vp],
    "lcasr_subsample_dwconv_bwd_weight": [vp, vp, i32, i64, i32, i32, vp, vp, vp],
    "lcasr_subsample_conv0_bwd": [vp, vp, vp, vp, i32, i32, i64, i32, vp, vp, vp],
    "lcasr_model_create": [C.POINTER(LcasrConfig), C.POINTER(LcasrWeights), C.POINTER(vp)],
    "lcasr_model_set_impl": [vp, i32, i32],
    "lcasr_model_set_attention_tail": [vp, i32, i32],
    "lcasr_attention_tail_plan": [i32, i64, i32, i32, vp, vp],
    "lcasr_model_set_timing": [vp, i32],
    "lcasr_model_get_timing": [vp, vp, vp, i32],
    "lcasr_model_forward": [vp, vp, i32, i64, vp, vp, i32, vp, i64, vp],
    "lcasr_model_forward_lengths": [vp, vp, i32, i64, vp, vp, vp, i32, vp, i64, vp],
    "lcasr_model_transcribe_host": [vp, vp, i32, i64, vp, vp, vp, vp, i64, vp],
    "lcasr_ctc_loss_fwd_ws": [vp, i32, i64, i32, vp, i64, vp, vp, i32, vp, vp, vp, i64, vp],
    "lcasr_ctc_loss_bwd_ws": [vp, i32, i64, i32, vp, i64, vp, vp, i32, vp, vp, vp, vp, vp, vp, i64, vp],
    "lcasr_ctc_loss_fwd_ab_ws": [vp, i32, i64, i32, vp, i64, vp, vp, i32, vp, vp, vp, vp, i64, vp],
    "lcasr_attention_partial": [vp, vp, vp, i32, i64, i64, i32, i32, vp, vp, vp],
    "lcasr_attention_merge": [vp, vp, i32, i64, i32, i32, vp, i32, vp],
    "lcasr_comm_unique_id": [C.c_char_p, C.c_char_p],
    "lcasr_comm_create": [C.c_char_p, i32, i32, C.c_char_p, C.POINTER(vp)],
    "lcasr_model_seqpar_block": [vp, i32, i32, i64, C.POINTER(i64), C.POINTER(i64)],
    "lcasr_model_forward_seqpar": [vp, vp, vp, i64, vp, vp, i32, vp, i64, vp],
    "lcasr_model_forward_seqpar_emulated": [vp, i32, vp, i64, vp, vp, i32, vp, i64, vp],
}
_OTHER = {
    "lcasr_abi_version": ([], i32),
    "lcasr_opt_chunk_elems": ([], i32),
    "lcasr_last_error": ([], C.c_char_p),
    "lcasr_launch_count": ([], i64),
    "lcasr_reset_launch_count": ([], None),
    "lcasr_out_length": ([i64], i64),
    "lcasr_melspec_frames": ([i64], i64),
    "lcasr_model_destroy": ([vp], None),
    "lcasr_model_workspace_bytes": ([vp, i32, i64], i64),
    "lcasr_model_transcribe_workspace_bytes": ([vp, i32, i64], i64),
    "lcasr_comm_destroy": ([vp], None),
    "lcasr_ctc_workspace_bytes": ([i32, i64, i64, i32], i64),
    "lcasr_attention_bwd_flash_workspace_bytes": ([i32, i64, i32], i64),
    "lcasr_model_seqpar_workspace_bytes": ([vp, i32, i32, i64], i64),
    "lcasr_model_seqpar_emulated_workspace_bytes": ([vp, i32, i64], i64),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES) + tuple(_OTHER)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C long-context-asr_b200/csrc`). lcasr_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, i32
    for name, (args, res) in _OTHER.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    if lib.lcasr_abi_version() != ABI_VERSION:
        raise ImportError(f"liblcasr_b200.so ABI {lib.lcasr_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    return lib


lib = _load()


class LcasrError(RuntimeError):
    pass


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib.lcasr_last_error().decode(errors="replace")
        raise LcasrError(f"lcasr_b200 {what} failed (status {status}): {msg}")


# Parameters can change behind PyTorch's version counters: the MADGRAD / BatchRenorm kernels write through raw pointers and
# `p.data.mul_()` (the reference's optimizer) does not bump `p._version` either.  Everything that may have changed weights
# (a training-mode forward, an optimizer step of this package) bumps this epoch; SCConformerXL re-packs its eval weights
# when it moved.  Callers that edit `.data` by hand (an EMA swap) call `model.invalidate_packed_weights()`.
WEIGHT_EPOCH = [0]


def bump_weight_epoch() -> None:
    WEIGHT_EPOCH[0] += 1


# Optional per-entry-point device timing (bench.py's kernel breakdown of the Python-orchestrated training step):
# set TIMING = {} to collect {name: [(start_event, end_event), ...]} on the current stream; None = off.
TIMING = None
TIMING_TAGS = False  # True: split entries by the shape tag some wrappers pass (per-shape GEMM table)


def call(name: str, *args, tag: str = "") -> None:
    if TIMING is None:
        check(getattr(lib, name)(*args), name)
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib, name)(*args), name)
    e1.record()
    TIMING.setdefault(name + tag if TIMING_TAGS else name, []).append((e0, e1))


def timing_summary(rec) -> dict:
    """{name: (total ms, calls)} from a TIMING record (call after a synchronize)."""
    return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in rec.items()}


def ptr(t):
    """device/host pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def dtype_code(torch_dtype) -> int:
    import torch
    if torch_dtype == torch.float32:
        return F32
    if torch_dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"lcasr_b200 supports float32 / bfloat16 tensors, got {torch_dtype}")


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
