"""Drop-in for ``lcasr/models/sconformer_xl.py``: same constructor kwargs, parameter / buffer names
and shapes (strict ``load_state_dict`` of a reference checkpoint works), same
``forward(audio_signal, length=None, cached_kvs=None, cached_kv_lengths=None, return_logits=False)``
returning ``{'final_posteriors', 'length'}`` — but the arithmetic is ONE call into the hand-written
sm_100a kernels behind the C ABI (``lcasr_model_forward``).  The sub-modules below only own
parameters; none of them has a PyTorch compute path (there is no fallback by design).

Reference map
  SCConformerXL.__init__ / forward      lcasr/models/sconformer_xl.py:31-252
  ConformerLayer                        lcasr/models/sconformer_xl.py:255-372
  ConvSubsampling (dw_striding)         lcasr/components/subsampling.py:165-428
  Attention                             lcasr/components/attention.py:448-551
  RotaryPositionalEmbedding             lcasr/components/rotary_emb.py:4-57
  ConformerConvolution / BatchRenorm1d  lcasr/components/convolution.py:41-124, batchrenorm.py:8-97
  FusedMLP                              lcasr/components/fused_dense.py:425-470
  PreNorm / Scale                       lcasr/components/wrappers.py:5-28
  ASRLinearSCDecoder                    lcasr/components/decoder.py:6-32
  BaseModel helpers                     lcasr/models/base.py:9-68
"""
from __future__ import annotations

import ctypes as C
import math
import warnings
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib as L

__all__ = ["SCConformerXL", "ConformerLayer", "RMSNorm", "BatchRenorm1d"]


class _ParamOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} only owns parameters; the computation is fused into "
                           "SCConformerXL.forward (lcasr_b200 has no PyTorch compute path)")


class RMSNorm(_ParamOnly):
    """lcasr/components/normalisation.py:6-47 — parameter is called ``scale``; eps=1e-8 outside the sqrt."""

    def __init__(self, d_model, p=-1.0, eps=1e-8, bias=False):
        super().__init__()
        if bias or not (p < 0.0 or p > 1.0):
            raise NotImplementedError("partial / biased RMSNorm is not on the hot path")
        self.eps, self.d = eps, d_model
        self.scale = nn.Parameter(torch.ones(d_model))


class LayerNorm(nn.LayerNorm):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("LayerNorm is fused into SCConformerXL.forward")


class BatchRenorm1d(_ParamOnly):
    """lcasr/components/batchrenorm.py:8-97 (eval path only on the inference hot path)."""

    def __init__(self, num_features, eps=1e-3, momentum=0.01):
        super().__init__()
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_std", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.eps, self.momentum = eps, momentum


class RotaryPositionalEmbedding(_ParamOnly):
    """lcasr/components/rotary_emb.py:4-57 — tables are rebuilt on device by lcasr_rope_table."""

    def __init__(self, dim, base=10000, learned_freq=False, rotary_interpolation_factor=1.0):
        super().__init__()
        if learned_freq:
            raise NotImplementedError("learned_rotary=True is not used by any released config")
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.learned_freq, self.dim = learned_freq, dim
        self.register_buffer("inv_freq", inv_freq)
        self.register_buffer("rotary_interpolation_factor", torch.tensor(rotary_interpolation_factor))

    def reset_if_needed(self):
        pass


class FusedMLP(_ParamOnly):
    """lcasr/components/fused_dense.py:425-470: fc1 -> gelu(tanh) -> fc2, hidden = 4*d."""

    def __init__(self, in_features, bias1=True, bias2=True, checkpoint_lvl=0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, in_features * 4, bias=bias1)
        self.fc2 = nn.Linear(in_features * 4, in_features, bias=bias2)


class PreNorm(_ParamOnly):
    def __init__(self, d_model, fn, norm, sandwich_norm=False):
        super().__init__()
        self.norm = norm(d_model)
        self.fn = fn


class Scale(_ParamOnly):
    def __init__(self, scale, fn):
        super().__init__()
        self.scale, self.fn = scale, fn


class Attention(_ParamOnly):
    """lcasr/components/attention.py:448-487 (parameters); qkv rows are ordered (h, dh, {q,k,v})."""

    def __init__(self, n_feats, head_dim, n_heads, **kwargs):
        super().__init__()
        # local attention (attention.py:321-328,466): query i sees keys [i - left, i + right]; -1 = unlimited
        def _win(direction):
            v = kwargs.get(f"attention_window_size_{direction}", None)
            return int(v) if v is not None else int(kwargs.get("attention_window_size", -1))
        self.left_window, self.right_window = _win("left"), _win("right")
        if kwargs.get("causal", False) or kwargs.get("qkv_bias", False) or kwargs.get("bias", False):
            raise NotImplementedError("causal / biased attention is not used by any released config")
        self.layer_idx = kwargs.get("layer_idx", None)
        self.n_feats, self.head_dim, self.n_heads = n_feats, head_dim, n_heads
        self.qkv_proj = nn.Linear(n_feats, 3 * n_heads * head_dim, bias=False)
        self.out_proj = nn.Linear(n_heads * head_dim, n_feats, bias=False)


class ConformerConvolution(_ParamOnly):
    """lcasr/components/convolution.py:41-101 (parameters)."""

    def __init__(self, d_model, kernel_size, norm_type="batch_renorm", exp_factor=1):
        super().__init__()
        if norm_type != "batch_renorm" or exp_factor != 1:
            raise NotImplementedError("conv_norm != 'batch_renorm' / conv_expansion_factor != 1: not on the hot path")
        assert (kernel_size - 1) % 2 == 0
        self.pointwise_conv1 = nn.Conv1d(d_model, 2 * d_model, 1)
        self.depthwise_conv = nn.Conv1d(d_model, d_model, kernel_size, padding=(kernel_size - 1) // 2, groups=d_model)
        self.batch_norm = BatchRenorm1d(d_model)
        self.pointwise_conv2 = nn.Conv1d(d_model, d_model, 1)


class ConformerLayer(_ParamOnly):
    """lcasr/models/sconformer_xl.py:255-343 (parameters)."""

    def __init__(self, d_model, conv_kernel_size, head_dim, n_heads, default_norm, bias_in_ff=False, layer_idx=0,
                 conv_expansion_factor=1, **kwargs):
        super().__init__()
        self.conv = PreNorm(d_model, ConformerConvolution(d_model, conv_kernel_size, kwargs.get("conv_norm", "batch_renorm"),
                                                          conv_expansion_factor), default_norm)
        self.ff1 = Scale(0.5, PreNorm(d_model, FusedMLP(d_model, bias1=bias_in_ff, bias2=bias_in_ff), default_norm))
        self.ff2 = Scale(0.5, PreNorm(d_model, FusedMLP(d_model, bias1=bias_in_ff, bias2=bias_in_ff), default_norm))
        self.attend = PreNorm(d_model, Attention(d_model, head_dim, n_heads, layer_idx=layer_idx, **kwargs), default_norm)
        self.norm_out = default_norm(d_model)


class ConvSubsampling(_ParamOnly):
    """lcasr/components/subsampling.py:250-323,374: Sequential indices 0|2,3|5,6 carry weights."""

    def __init__(self, subsampling_factor, feat_in, feat_out, conv_channels):
        super().__init__()
        if subsampling_factor != 8:
            raise NotImplementedError("only 8x dw_striding subsampling is on the hot path")
        self.subsampling_factor = subsampling_factor
        Cc = conv_channels
        self.conv = nn.Sequential(
            nn.Conv2d(1, Cc, 3, 2, 1), nn.SiLU(),
            nn.Conv2d(Cc, Cc, 3, 2, 1, groups=Cc), nn.Conv2d(Cc, Cc, 1), nn.SiLU(),
            nn.Conv2d(Cc, Cc, 3, 2, 1, groups=Cc), nn.Conv2d(Cc, Cc, 1), nn.SiLU())
        f = feat_in
        for _ in range(3):
            f = (f - 1) // 2 + 1
        self.feat_sub = f
        self.out = nn.Linear(Cc * f, feat_out, bias=False)


class ASRLinearSCDecoder(_ParamOnly):
    """lcasr/components/decoder.py:6-32 (parameters)."""

    def __init__(self, d_model, vocab_size, norm=False, norm_fn=LayerNorm):
        super().__init__()
        self.num_classes = vocab_size + 1
        self.ff = nn.Linear(d_model, self.num_classes)
        self.reprojection = nn.Linear(self.num_classes, d_model)
        self.norm = norm_fn(d_model) if norm else nn.Identity()


_COMPUTE_DTYPES = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp32": torch.float32, "float32": torch.float32}


def pack_qkv_rotary_interleaved(qkv: torch.Tensor, H: int, Dh: int) -> torch.Tensor:
    """Weight rows for the qkv GEMM with the rotary embedding in its epilogue (csrc/gemm_tc.cu, TG_EPI_ROPE).
    qkv: [3*H*Dh, d] with rows already in [q | k | v] order.  Inside every q / k head the rows are interleaved —
    new row 2i <- old row i, new row 2i+1 <- old row i + Dh/2 — so that a ``rotate_half`` pair (rotary_emb.py:61-66) is two
    ADJACENT output columns of one epilogue thread.  q.k is invariant under the common permutation of the head dimension,
    so attention needs no un-permute; v is left alone."""
    d = qkv.shape[1]
    qk = qkv[: 2 * H * Dh].reshape(2 * H, Dh, d)
    qk_il = torch.stack([qk[:, : Dh // 2], qk[:, Dh // 2:]], dim=2).reshape(2 * H * Dh, d)
    return torch.cat([qk_il, qkv[2 * H * Dh:]], 0)


def pack_glu_blocks(w1: torch.Tensor, b1: torch.Tensor):
    """pointwise_conv1 rows for the GEMM with ``F.glu`` in its epilogue (TG_EPI_GLU): w1 [2d, d] = [value rows | gate rows]
    (convolution.py:107-108) becomes 64-row blocks of 32 value channels followed by THEIR 32 gate channels, so that one
    epilogue warp holds a (value, gate) chunk pair; the bias likewise."""
    d = w1.shape[0] // 2
    w = torch.stack([w1[:d].reshape(d // 32, 32, -1), w1[d:].reshape(d // 32, 32, -1)], 1).reshape(2 * d, -1)
    b = torch.stack([b1[:d].reshape(d // 32, 32), b1[d:].reshape(d // 32, 32)], 1).reshape(2 * d)
    return w, b


class SCConformerXL(nn.Module):
    """Same signature as the reference (sconformer_xl.py:32-64).  One extra, optional kwarg:
    ``compute_dtype`` ('bf16' default: tcgen05 tensor-core path; 'fp32': SIMT fp32 parity mode)."""

    def __init__(self, vocab_size=128, feat_in=80, subsampling="dw_striding", subsampling_factor=8,
                 subsampling_conv_channels=256, subsampling_act="silu", subsampling_norm_out=False, n_layers=6,
                 d_model=768, n_heads=6, head_dim=128, expansion_factor=4, dropout_ff=0.0, dropout_conv=0.0,
                 dropout_attn=0.0, checkpoint_every_n_layers=0, conv_kernel_size=9, conv_expansion_factor=1,
                 decoder_norm=False, use_rotary=False, rotary_interpolation_factor=1.0, learned_rotary=False,
                 fourier_pos_enc=False, self_conditioning=True, default_norm="layer_norm", sandwich_norm=False,
                 bias_in_ff=False, transformer=False, legasee_double_norm=True, **kwargs):
        super().__init__()
        assert subsampling_act in ("silu", "relu", "gelu", "none"), f"subsampling_act must be one of silu/relu/gelu/none (got {subsampling_act})"
        assert default_norm in ("rms_norm", "layer_norm"), f"default_norm must be rms_norm or layer_norm (got {default_norm})"
        unsupported = []
        if subsampling != "dw_striding": unsupported.append(f"subsampling={subsampling!r}")
        if subsampling_act != "silu": unsupported.append(f"subsampling_act={subsampling_act!r}")
        if subsampling_norm_out: unsupported.append("subsampling_norm_out=True")
        if fourier_pos_enc: unsupported.append("fourier_pos_enc=True")
        if sandwich_norm: unsupported.append("sandwich_norm=True")
        if transformer: unsupported.append("transformer=True")
        if dropout_ff or dropout_conv or dropout_attn: unsupported.append("dropout > 0")
        if unsupported:
            raise NotImplementedError("lcasr_b200 covers the released configs only; not on the hot path: " + ", ".join(unsupported))

        self.feat_in, self.n_layers, self.d_model, self.n_heads, self.head_dim = feat_in, n_layers, d_model, n_heads, head_dim
        self.expansion_factor, self.conv_kernel_size, self.conv_expansion_factor = expansion_factor, conv_kernel_size, conv_expansion_factor
        self.rotary_interpolation_factor, self.learned_rotary = rotary_interpolation_factor, learned_rotary
        self.self_conditioning, self.sandwich_norm, self.bias_in_ff, self.transformer = self_conditioning, sandwich_norm, bias_in_ff, transformer
        self.legasee_double_norm = legasee_double_norm
        self.checkpoint_subsampling = kwargs.get("checkpoint_subsampling", False)  # accepted, irrelevant (no autograd graph)
        self.flash_attn = kwargs.get("flash_attn", True)
        self.checkpoint_every_n_layers = checkpoint_every_n_layers
        self.dropout_ff, self.dropout_conv, self.dropout_attn = dropout_ff, dropout_conv, dropout_attn
        self.subsampling_mode, self.subsampling_factor = subsampling, subsampling_factor
        self.subsampling_conv_channels = subsampling_conv_channels if subsampling_conv_channels != -1 else d_model
        self.decoder_norm, self.use_rotary = decoder_norm, use_rotary
        self.default_norm_name = default_norm
        cd = kwargs.pop("compute_dtype", "bf16")
        self.compute_dtype = _COMPUTE_DTYPES[cd] if isinstance(cd, str) else cd
        norm_cls = RMSNorm if default_norm == "rms_norm" else LayerNorm

        self.whitelist_weight_decay_modules = (nn.LayerNorm, RMSNorm, LayerNorm, BatchRenorm1d, nn.GroupNorm)
        self.blacklist_weight_decay_modules = (nn.Linear, FusedMLP, nn.Conv1d, nn.Conv2d, RotaryPositionalEmbedding)

        self.rotary_pos_emb = None
        if use_rotary:
            self.rotary_pos_emb = RotaryPositionalEmbedding(head_dim, kwargs.get("rotary_base_freq", 10000), learned_rotary,
                                                            rotary_interpolation_factor)
        self.fourier_pos_enc = nn.Identity()
        self.decoder = ASRLinearSCDecoder(d_model, vocab_size, norm=decoder_norm, norm_fn=norm_cls)
        self.subsampling = ConvSubsampling(subsampling_factor, feat_in, d_model, self.subsampling_conv_channels)
        layer_kwargs = {k: v for k, v in kwargs.items() if k not in ("layer_idx",)}
        self.layers = nn.ModuleList([
            ConformerLayer(d_model=d_model, conv_kernel_size=conv_kernel_size, head_dim=head_dim, n_heads=n_heads,
                           default_norm=norm_cls, bias_in_ff=bias_in_ff, layer_idx=i,
                           conv_expansion_factor=conv_expansion_factor, **layer_kwargs) for i in range(n_layers)])

        self._handle: Optional[int] = None
        self._pack_key = None
        self._packed: Dict[str, torch.Tensor] = {}
        self._keepalive = None
        self._workspace: Optional[torch.Tensor] = None
        self._impl = (L.GEMM_AUTO, L.ATTN_AUTO)
        self.last_argmax: Optional[torch.Tensor] = None
        # opt-in: replay the ~135 launches of an equal-length eval forward as ONE CUDA graph per (B, T) — for
        # launch-bound shapes (10-s contexts: 1.97 ms of launches for ~0.4 ms of kernels)
        self.cuda_graphs = False
        # eval() mode WITH gradients (test-time adaptation, lcasr/eval/dynamic_eval.py:47-100): opt-in, because a plain
        # model(x) in eval mode outside torch.no_grad() should keep taking the fused inference call
        self.grad_in_eval = False
        self._graphs: Dict = {}
        self.graph_launches_replayed = 0  # kernels launched through graph replays (they bypass lcasr_launch_count)

    # ---- BaseModel API (lcasr/models/base.py) -------------------------------------------------
    def print_total_params(self, only_trainable=False):
        total = sum(p.numel() for p in self.parameters() if p.requires_grad or not only_trainable)
        print(("Total trainable params: " if only_trainable else "Total params: ") + ": ", total / 1e6, "M")
        return total

    @staticmethod
    def create_custom_forward(module):
        def custom_forward(*args, **kwargs):
            return module(*args, **kwargs)
        return custom_forward

    def get_param_groups(self, optim_args):
        """base.py:25-68 (note the reference's naming: 'blacklist' AND 'whitelist' modules both end up
        undecayed/decayed exactly as written there)."""
        wd = optim_args.get("weight_decay", 0.0)
        if wd <= 0.0:
            return self.parameters()
        decay, no_decay = set(), set()
        for mn, m in self.named_modules():
            for pn, _ in m.named_parameters():
                fpn = f"{mn}.{pn}" if mn else pn
                if pn.endswith("bias"):
                    no_decay.add(fpn)
                elif isinstance(m, self.blacklist_weight_decay_modules):
                    no_decay.add(fpn)
                elif isinstance(m, self.whitelist_weight_decay_modules):
                    decay.add(fpn)
        param_dict = dict(self.named_parameters())
        assert not (decay & no_decay), f"parameters {decay & no_decay} made it into both decay/no_decay sets!"
        assert not (param_dict.keys() - (decay | no_decay)), "parameters were not separated into either decay/no_decay set!"
        return [{"params": [param_dict[pn] for pn in sorted(decay)], "weight_decay": wd},
                {"params": [param_dict[pn] for pn in sorted(no_decay)], "weight_decay": 0.0}]

    # ---- weight packing (done once per state_dict version; plumbing, not the hot path) ---------
    def set_kernel_impl(self, gemm: int = L.GEMM_AUTO, attn: int = L.ATTN_AUTO):
        """test hook: force SIMT / tcgen05 kernels."""
        self._impl = (gemm, attn)
        if self._handle is not None:
            L.call("lcasr_model_set_impl", self._handle, gemm, attn)

    def set_attention_tail(self, tail_pairs: int = -1, key_pieces: int = 0):
        """tuning hook (lcasr_model_set_attention_tail): -1 = planned automatically, 0 = off, > 0 = forced split."""
        self._attn_tail = (int(tail_pairs), int(key_pieces))
        if self._handle is not None:
            L.call("lcasr_model_set_attention_tail", self._handle, *self._attn_tail)

    def _state_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers())) + \
            (self.compute_dtype, L.WEIGHT_EPOCH[0])

    def invalidate_packed_weights(self):
        """call after editing parameters through `.data` / raw pointers outside this package (version counters do not move)"""
        self._pack_key = None

    def _destroy(self):
        if self._handle is not None:
            L.lib.lcasr_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _build(self, device):
        self._destroy()
        self._graphs = {}  # captured graphs hold the old handle / weight pointers
        cdt = self.compute_dtype
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        H, Dh, d, Cc = self.n_heads, self.head_dim, self.d_model, self.subsampling_conv_channels
        F3 = self.subsampling.feat_sub
        P: Dict[str, torch.Tensor] = {}

        def vec(key):
            t = sd.get(key)
            return None if t is None else t.to(device=device, dtype=torch.float32).contiguous()

        def mat(t):
            return t.to(device=device, dtype=cdt).contiguous()

        rms = self.default_norm_name == "rms_norm"

        def norm_wb(prefix):
            return (vec(prefix + ".scale"), None) if rms else (vec(prefix + ".weight"), vec(prefix + ".bias"))

        w = L.LcasrWeights()
        layers = (L.LcasrLayerWeights * self.n_layers)()
        keep = []
        packed: Dict[str, Optional[torch.Tensor]] = {}  # the same tensors by name (sequence-parallel driver, tests)

        cur_prefix = [""]  # "" for model-level tensors, "layers.<l>." inside the layer loop

        def put(struct, field, t):
            if t is not None:
                keep.append(t)
            packed[cur_prefix[0] + field] = t
            setattr(struct, field, L.ptr(t))

        put(w, "conv0_w", vec("subsampling.conv.0.weight").reshape(Cc, 9).contiguous())
        put(w, "conv0_b", vec("subsampling.conv.0.bias"))
        for name, i in (("1", 2), ("2", 5)):
            put(w, f"dw{name}_w", vec(f"subsampling.conv.{i}.weight").reshape(Cc, 9).contiguous())
            put(w, f"dw{name}_b", vec(f"subsampling.conv.{i}.bias"))
            put(w, f"pw{name}_w", mat(sd[f"subsampling.conv.{i + 1}.weight"].reshape(Cc, Cc)))
            put(w, f"pw{name}_b", vec(f"subsampling.conv.{i + 1}.bias"))
        # reference feature index is c*F3+f (subsampling.py:422-423); our activations are channels-last (f*C+c)
        so = sd["subsampling.out.weight"].reshape(d, Cc, F3).permute(0, 2, 1).reshape(d, F3 * Cc)
        put(w, "sub_out_w", mat(so))
        put(w, "inv_freq", vec("rotary_pos_emb.inv_freq") if self.use_rotary else None)
        if self.decoder_norm:
            nw, nb = norm_wb("decoder.norm")
            put(w, "dec_norm_w", nw); put(w, "dec_norm_b", nb)
        put(w, "dec_ff_w", mat(sd["decoder.ff.weight"])); put(w, "dec_ff_b", vec("decoder.ff.bias"))
        put(w, "dec_rep_w", mat(sd["decoder.reprojection.weight"])); put(w, "dec_rep_b", vec("decoder.reprojection.bias"))

        for l in range(self.n_layers):
            p, lw = f"layers.{l}.", layers[l]
            cur_prefix[0] = p
            for ff in ("ff1", "ff2"):
                nw, nb = norm_wb(p + ff + ".fn.norm")
                put(lw, ff + "_norm_w", nw); put(lw, ff + "_norm_b", nb)
                put(lw, ff + "_fc1_w", mat(sd[p + ff + ".fn.fn.fc1.weight"])); put(lw, ff + "_fc1_b", vec(p + ff + ".fn.fn.fc1.bias"))
                put(lw, ff + "_fc2_w", mat(sd[p + ff + ".fn.fn.fc2.weight"])); put(lw, ff + "_fc2_b", vec(p + ff + ".fn.fn.fc2.bias"))
            nw, nb = norm_wb(p + "attend.norm")
            put(lw, "attn_norm_w", nw); put(lw, "attn_norm_b", nb)
            # attention.py:485: rows ordered (h, dh, qkv) -> de-interleave once to [q | k | v]
            qkv = sd[p + "attend.fn.qkv_proj.weight"].reshape(H, Dh, 3, d).permute(2, 0, 1, 3).reshape(3 * H * Dh, d)
            put(lw, "qkv_w", mat(qkv)); put(lw, "out_w", mat(sd[p + "attend.fn.out_proj.weight"]))
            # fused-rotary form: inside every q / k head interleave the two rotate_half halves (rotary_emb.py:61-66 pairs
            # (i, i + Dh/2)) so that a rotation pair is two adjacent output columns of the qkv GEMM
            put(lw, "qkv_w_il", mat(pack_qkv_rotary_interleaved(qkv, H, Dh)) if cdt == torch.bfloat16 and self.use_rotary else None)
            nw, nb = norm_wb(p + "conv.norm")
            put(lw, "conv_norm_w", nw); put(lw, "conv_norm_b", nb)
            put(lw, "pw1_w", mat(sd[p + "conv.fn.pointwise_conv1.weight"].reshape(2 * d, d))); put(lw, "pw1_b", vec(p + "conv.fn.pointwise_conv1.bias"))
            if cdt == torch.bfloat16 and d % 32 == 0:  # fused-GLU form: 64-row blocks [32 value channels | their 32 gate channels]
                w1 = sd[p + "conv.fn.pointwise_conv1.weight"].reshape(2 * d, d)
                b1 = sd[p + "conv.fn.pointwise_conv1.bias"]
                w1g, b1g = pack_glu_blocks(w1, b1)
                put(lw, "pw1_w_glu", mat(w1g))
                put(lw, "pw1_b_glu", b1g.to(device=device, dtype=torch.float32).contiguous())
            else:
                put(lw, "pw1_w_glu", None); put(lw, "pw1_b_glu", None)
            put(lw, "dw_w", vec(p + "conv.fn.depthwise_conv.weight").reshape(d, self.conv_kernel_size).contiguous())
            put(lw, "dw_b", vec(p + "conv.fn.depthwise_conv.bias"))
            put(lw, "brn_mean", vec(p + "conv.fn.batch_norm.running_mean")); put(lw, "brn_std", vec(p + "conv.fn.batch_norm.running_std"))
            put(lw, "brn_w", vec(p + "conv.fn.batch_norm.weight")); put(lw, "brn_b", vec(p + "conv.fn.batch_norm.bias"))
            put(lw, "pw2_w", mat(sd[p + "conv.fn.pointwise_conv2.weight"].reshape(d, d))); put(lw, "pw2_b", vec(p + "conv.fn.pointwise_conv2.bias"))
            nw, nb = norm_wb(p + "norm_out")
            put(lw, "norm_out_w", nw); put(lw, "norm_out_b", nb)
        w.layers_host = C.cast(layers, C.POINTER(L.LcasrLayerWeights))

        cfg = L.LcasrConfig(
            abi_version=L.ABI_VERSION, n_layers=self.n_layers, d_model=d, n_heads=H, head_dim=Dh, feat_in=self.feat_in,
            conv_channels=Cc, conv_kernel_size=self.conv_kernel_size, num_classes=self.decoder.num_classes,
            norm_kind=L.NORM_RMSNORM if rms else L.NORM_LAYERNORM, decoder_norm=int(self.decoder_norm),
            use_rotary=int(self.use_rotary), self_conditioning=int(self.self_conditioning),
            legasee_double_norm=int(self.legasee_double_norm), bias_in_ff=int(self.bias_in_ff),
            compute_dtype=L.dtype_code(cdt),
            rotary_interp=float(sd["rotary_pos_emb.rotary_interpolation_factor"]) if self.use_rotary else 1.0,
            norm_eps=1e-8 if rms else 1e-5,
            attn_window_left=self.layers[0].attend.fn.left_window, attn_window_right=self.layers[0].attend.fn.right_window)
        handle = L.vp()
        with torch.cuda.device(device):  # the handle owns side streams: they must live on the model's device
            L.call("lcasr_model_create", C.byref(cfg), C.byref(w), C.byref(handle))
        self._handle = handle.value
        self._keepalive = (keep, layers, w)
        self._packed = packed
        L.call("lcasr_model_set_impl", self._handle, *self._impl)
        L.call("lcasr_model_set_attention_tail", self._handle, *getattr(self, "_attn_tail", (-1, 0)))

    def _ensure_built(self, device):
        key = (self._state_key(), str(device))
        if self._handle is None or key != self._pack_key:
            self._build(device)
            self._pack_key = key

    def _ws(self, nbytes: int, device) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != device:
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._workspace

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, audio_signal, length=None, cached_kvs=None, cached_kv_lengths=None, return_logits=False):
        """Same signature and result as the reference (sconformer_xl.py:162-252).  ``model.eval()``: one fused
        inference call (no autograd graph).  ``model.train()``: the training path (training.py) — BatchRenorm uses
        batch statistics and updates its running buffers, and ``final_posteriors`` carries a grad_fn whose
        backward is the hand-written backward pass, so ``loss.backward()`` fills ``.grad`` of every parameter.
        ``model.eval()`` with ``model.grad_in_eval = True`` and gradients enabled: the same autograd node with
        BatchRenorm in its eval form (running statistics, buffers untouched) — dynamic_eval.py's adaptation step."""
        if self.training or (self.grad_in_eval and torch.is_grad_enabled()):
            return self._forward_train(audio_signal, length, cached_kvs, cached_kv_lengths, return_logits)
        return self._forward_eval(audio_signal, length, cached_kvs, cached_kv_lengths, return_logits)

    def _forward_train(self, audio_signal, length, cached_kvs, cached_kv_lengths, return_logits):
        from .training import train_forward
        if cached_kvs is not None or cached_kv_lengths is not None:
            raise NotImplementedError("cached_kvs is dead code in the reference (SURVEY §3.2) and not supported")
        if not audio_signal.is_cuda:
            raise RuntimeError("lcasr_b200.SCConformerXL runs on CUDA (sm_100a) only; there is no CPU fallback")
        if audio_signal.dim() != 3 or audio_signal.shape[1] != self.feat_in:
            raise ValueError(f"audio_signal must be [B, {self.feat_in}, T], got {tuple(audio_signal.shape)}")
        if self.compute_dtype != torch.bfloat16:
            raise NotImplementedError("the training path computes in bf16 (the reference trains under bf16 autocast)")
        if self.layers[0].attend.fn.left_window >= 0 or self.layers[0].attend.fn.right_window >= 0:
            raise NotImplementedError("windowed attention is an evaluation mode (eval/run.py:38-43); training uses full attention")
        B, _, T = audio_signal.shape
        tok_lens = None
        if length is not None:  # padded batch (exp/train.py:236-241): token counts on the host, like the reference's
            lens = [int(v) for v in (length.tolist() if torch.is_tensor(length) else length)]  # length.max() == length.min()
            if len(lens) != B or any(v < 1 or v > T for v in lens):
                raise ValueError(f"length must hold {B} frame counts in [1, {T}], got {lens}")
            tok_lens = [int(L.lib.lcasr_out_length(v)) for v in lens]
            if max(tok_lens) != int(L.lib.lcasr_out_length(T)):  # same rule as the evaluation path (shape clash in the reference)
                raise ValueError("the longest recording must span the padded batch (length.max() == T up to the 8x rounding)")
        masked = tok_lens if (tok_lens is not None and min(tok_lens) != max(tok_lens)) else None  # sconformer_xl.py:204-205
        L.bump_weight_epoch()  # a training / adaptation step follows: the packed eval weights are stale from here on
        lp = train_forward(self, audio_signal.detach().to(torch.float32).contiguous(), masked, brn_eval=not self.training,
                           return_logits=bool(return_logits))
        N = lp.shape[1]
        out_len = torch.full((B,), N, dtype=torch.int32, device=audio_signal.device) if tok_lens is None else \
            torch.tensor(tok_lens, dtype=torch.int32, device=audio_signal.device)
        return {"final_posteriors": lp, "length": out_len}

    @torch.no_grad()
    def _forward_eval(self, audio_signal, length=None, cached_kvs=None, cached_kv_lengths=None, return_logits=False):
        """audio_signal [B, feat_in, T] fp32 CUDA; length [B] frames per recording, or None (= T).
        Ragged batches take the key-padding-mask path (sconformer_xl.py:204-215): rows of padded tokens
        in `final_posteriors` are unspecified, exactly `length[b]` rows of recording b are meaningful."""
        if cached_kvs is not None or cached_kv_lengths is not None:
            raise NotImplementedError("cached_kvs is dead code in the reference (SURVEY §3.2) and not supported")
        if not audio_signal.is_cuda:
            raise RuntimeError("lcasr_b200.SCConformerXL runs on CUDA (sm_100a) only; there is no CPU fallback")
        if audio_signal.dim() != 3 or audio_signal.shape[1] != self.feat_in:
            raise ValueError(f"audio_signal must be [B, {self.feat_in}, T], got {tuple(audio_signal.shape)}")
        B, _, T = audio_signal.shape
        device = audio_signal.device
        N = int(L.lib.lcasr_out_length(T))
        tok_len = None  # None = every recording fills the batch (sconformer_xl.py:204-205)
        if length is not None:
            lens = [int(v) for v in (length.tolist() if torch.is_tensor(length) else length)]
            if len(lens) != B or any(v < 1 or v > T for v in lens):
                raise ValueError(f"length must hold {B} frame counts in [1, {T}], got {lens}")
            toks = [int(L.lib.lcasr_out_length(v)) for v in lens]  # subsampling.py:557-567 per recording
            if max(toks) != N:
                # the reference sizes its rotary tables and masks by length.max() (:188,198); a batch whose
                # longest recording is shorter than the padded tensor is a caller error there too (shape clash)
                raise ValueError("the longest recording must span the padded batch (length.max() == T up to the 8x rounding)")
            if min(toks) != max(toks):
                tok_len = torch.tensor(toks, dtype=torch.int32, device=device)
            length_out = torch.tensor(toks, dtype=torch.int32, device=device)
        else:
            length_out = torch.full((B,), N, dtype=torch.int32, device=device)
        x = audio_signal.to(torch.float32).contiguous()
        self._ensure_built(device)
        V1 = self.decoder.num_classes
        if self.cuda_graphs and tok_len is None:
            out, argmax = self._forward_graph(x, B, T, N, V1, return_logits, device)
            self.last_argmax = None if return_logits else argmax
            return {"final_posteriors": out, "length": length_out}
        out = torch.empty(B, N, V1, dtype=torch.float32, device=device)
        argmax = torch.empty(B, N, dtype=torch.int32, device=device)
        nbytes = int(L.lib.lcasr_model_workspace_bytes(self._handle, B, T))
        ws = self._ws(nbytes, device)
        with torch.cuda.device(device):
            L.call("lcasr_model_forward_lengths", self._handle, L.ptr(x), B, T, L.ptr(tok_len), L.ptr(out), L.ptr(argmax),
                   int(return_logits), L.ptr(ws), ws.numel(), L.current_stream())
        self.last_argmax = None if return_logits else argmax
        return {"final_posteriors": out, "length": length_out}

    def _forward_graph(self, x, B, T, N, V1, return_logits, device):
        """the C-ABI forward is allocation-free and asynchronous on one stream, so it captures as is: static input /
        output / workspace buffers per (B, T), one graph launch per call, results cloned out of the static buffers"""
        key = (B, T, bool(return_logits), str(device))
        ent = self._graphs.get(key)
        if ent is None:
            nbytes = int(L.lib.lcasr_model_workspace_bytes(self._handle, B, T))
            ent = dict(x=torch.empty_like(x), out=torch.empty(B, N, V1, dtype=torch.float32, device=device),
                       am=torch.empty(B, N, dtype=torch.int32, device=device), ws=torch.empty(nbytes, dtype=torch.uint8, device=device))

            def run():
                L.call("lcasr_model_forward_lengths", self._handle, L.ptr(ent["x"]), B, T, None, L.ptr(ent["out"]), L.ptr(ent["am"]),
                       int(return_logits), L.ptr(ent["ws"]), ent["ws"].numel(), L.current_stream())
            ent["x"].copy_(x)
            with torch.cuda.device(device):
                run()  # un-captured warm-up: function attributes, tensor-map entry point
                torch.cuda.synchronize(device)
                g = torch.cuda.CUDAGraph()
                n0 = int(L.lib.lcasr_launch_count())
                with torch.cuda.graph(g):
                    run()
                ent["launches"] = int(L.lib.lcasr_launch_count()) - n0
            ent["graph"] = g
            self._graphs[key] = ent
        ent["x"].copy_(x)
        ent["graph"].replay()
        self.graph_launches_replayed += ent["launches"]
        return ent["out"].clone(), ent["am"].clone()

    @torch.no_grad()
    def transcribe_host(self, spec_host: torch.Tensor):
        """End-to-end call used by bench.py's e2e leg: pinned host spectrogram [B,feat_in,T] in,
        greedy token ids out; H2D, forward, collapse and D2H all inside the C ABI."""
        assert not spec_host.is_cuda and spec_host.dtype == torch.float32 and spec_host.is_contiguous()
        B, _, T = spec_host.shape
        device = torch.device("cuda", torch.cuda.current_device())
        self._ensure_built(device)
        N = int(L.lib.lcasr_out_length(T))
        nbytes = int(L.lib.lcasr_model_transcribe_workspace_bytes(self._handle, B, T))
        ws = self._ws(nbytes, device)
        if getattr(self, "_logp_buf", None) is None or self._logp_buf.numel() < B * N * self.decoder.num_classes:
            self._logp_buf = torch.empty(B * N * self.decoder.num_classes, dtype=torch.float32, device=device)
            self._tok_host = torch.empty(B, N, dtype=torch.int32).pin_memory()
            self._cnt_host = torch.empty(B, dtype=torch.int32).pin_memory()
        if self._tok_host.shape != (B, N):
            self._tok_host = torch.empty(B, N, dtype=torch.int32).pin_memory()
            self._cnt_host = torch.empty(B, dtype=torch.int32).pin_memory()
        L.call("lcasr_model_transcribe_host", self._handle, spec_host.data_ptr(), B, T, self._tok_host.data_ptr(),
               self._cnt_host.data_ptr(), L.ptr(self._logp_buf), L.ptr(ws), ws.numel(), L.current_stream())
        return [self._tok_host[b, : int(self._cnt_host[b])].tolist() for b in range(B)]
