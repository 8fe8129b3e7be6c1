"""The training step of ``SCConformerXL`` (cfg 5 of BASELINE.json): train-mode forward that keeps what the
backward needs, and the hand-written backward, exposed to PyTorch as ONE autograd node so that the reference's
training loop (exp/train.py:236-262: ``out = model(...); loss = ctc(...); loss.backward(); optimizer.step()``)
runs unchanged.  Host orchestration is Python (the reference's host side is Python); every arithmetic step is a
kernel of liblcasr_b200.so called through the C ABI (train_ops.py / ops.py).  PyTorch allocates the buffers,
owns the stream and carries the gradient all-reduce (NCCL) of data-parallel training.

What the reference does with torch.autograd over its modules, layer by layer:
  ConvSubsampling          subsampling.py:277-323,384-428
  Scale(0.5, PreNorm(FusedMLP))   sconformer_xl.py:300-322, fused_dense.py:464-470
  PreNorm(Attention)       attention.py:483-551, rotary_emb.py:44-73
  PreNorm(ConformerConvolution)   convolution.py:103-124, batchrenorm.py:52-92 (TRAINING branch: batch statistics)
  norm_out, self-conditioning, decoder   sconformer_xl.py:241-247, decoder.py:17-32

Precision: bf16 GEMM operands / activations with fp32 accumulation (the reference trains under bf16 autocast,
exp/train.py:225); fp32 residual stream, LayerNorm statistics, softmax / log-softmax, BatchRenorm statistics,
residual-stream gradient and parameter gradients.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops
from . import train_ops as T

BF = torch.bfloat16


def _al4(n: int) -> int:
    return (n + 3) // 4 * 4


class TrainEngine:
    """Per-model helper: weight packing, gradient buffer layout, forward/backward orchestration."""

    def __init__(self, model):
        self.m = model
        self._rope: Dict = {}
        self._layout: Optional[Dict[str, tuple]] = None
        self.dp_group = None      # torch.distributed process group: all-reduce gradients inside backward
        self.dp_average = True    # divide by the world size (DistributedDataParallel semantics)
        self.dp_overlap = True    # per-layer all-reduces behind the rest of the backward (False: one all-reduce at the end)
        self.trace = None
        self.last_flat_grad: Optional[torch.Tensor] = None

    # ---- parameters -------------------------------------------------------------------------------------
    def pack(self, device) -> Dict[str, Optional[torch.Tensor]]:
        """bf16 GEMM operands (qkv rows de-interleaved, subsampling.out columns channels-last) + fp32 vectors,
        keyed like model._packed.  Runs every step (the optimizer just changed the weights); no host sync."""
        m = self.m
        sd = {k: v.detach() for k, v in m.state_dict(keep_vars=True).items()}
        H, Dh, d, Cc, F3 = m.n_heads, m.head_dim, m.d_model, m.subsampling_conv_channels, m.subsampling.feat_sub
        rms = m.default_norm_name == "rms_norm"
        P: Dict[str, Optional[torch.Tensor]] = {}
        names: Dict[str, str] = {}  # packed key -> parameter name

        def vec(pk, key, shape=None):
            t = sd.get(key)
            if t is not None:
                t = t.to(device=device, dtype=torch.float32)
                t = (t.reshape(shape) if shape is not None else t).contiguous()
                names[pk] = key
            P[pk] = t

        def mat(pk, key, t):
            P[pk] = t.to(device=device, dtype=BF).contiguous()
            names[pk] = key

        def norm(pk, prefix):
            if rms:
                vec(pk + "_w", prefix + ".scale"); P[pk + "_b"] = None
            else:
                vec(pk + "_w", prefix + ".weight"); vec(pk + "_b", prefix + ".bias")

        vec("conv0_w", "subsampling.conv.0.weight", (Cc, 9)); vec("conv0_b", "subsampling.conv.0.bias")
        for nm, i in (("1", 2), ("2", 5)):
            vec(f"dw{nm}_w", f"subsampling.conv.{i}.weight", (Cc, 9)); vec(f"dw{nm}_b", f"subsampling.conv.{i}.bias")
            mat(f"pw{nm}_w", f"subsampling.conv.{i + 1}.weight", sd[f"subsampling.conv.{i + 1}.weight"].reshape(Cc, Cc))
            vec(f"pw{nm}_b", f"subsampling.conv.{i + 1}.bias")
        mat("sub_out_w", "subsampling.out.weight",
            sd["subsampling.out.weight"].reshape(d, Cc, F3).permute(0, 2, 1).reshape(d, F3 * Cc))
        if m.decoder_norm:
            norm("dec_norm", "decoder.norm")
        mat("dec_ff_w", "decoder.ff.weight", sd["decoder.ff.weight"]); vec("dec_ff_b", "decoder.ff.bias")
        mat("dec_rep_w", "decoder.reprojection.weight", sd["decoder.reprojection.weight"])
        vec("dec_rep_b", "decoder.reprojection.bias")
        for l in range(m.n_layers):
            p, q = f"layers.{l}.", f"layers.{l}."
            for ff in ("ff1", "ff2"):
                norm(q + ff + "_norm", p + ff + ".fn.norm")
                mat(q + ff + "_fc1_w", p + ff + ".fn.fn.fc1.weight", sd[p + ff + ".fn.fn.fc1.weight"])
                mat(q + ff + "_fc2_w", p + ff + ".fn.fn.fc2.weight", sd[p + ff + ".fn.fn.fc2.weight"])
                vec(q + ff + "_fc1_b", p + ff + ".fn.fn.fc1.bias"); vec(q + ff + "_fc2_b", p + ff + ".fn.fn.fc2.bias")
            norm(q + "attn_norm", p + "attend.norm")
            mat(q + "qkv_w", p + "attend.fn.qkv_proj.weight",
                sd[p + "attend.fn.qkv_proj.weight"].reshape(H, Dh, 3, d).permute(2, 0, 1, 3).reshape(3 * H * Dh, d))
            mat(q + "out_w", p + "attend.fn.out_proj.weight", sd[p + "attend.fn.out_proj.weight"])
            norm(q + "conv_norm", p + "conv.norm")
            mat(q + "pw1_w", p + "conv.fn.pointwise_conv1.weight", sd[p + "conv.fn.pointwise_conv1.weight"].reshape(2 * d, d))
            vec(q + "pw1_b", p + "conv.fn.pointwise_conv1.bias")
            vec(q + "dw_w", p + "conv.fn.depthwise_conv.weight", (d, m.conv_kernel_size)); vec(q + "dw_b", p + "conv.fn.depthwise_conv.bias")
            vec(q + "brn_w", p + "conv.fn.batch_norm.weight"); vec(q + "brn_b", p + "conv.fn.batch_norm.bias")
            mat(q + "pw2_w", p + "conv.fn.pointwise_conv2.weight", sd[p + "conv.fn.pointwise_conv2.weight"].reshape(d, d))
            vec(q + "pw2_b", p + "conv.fn.pointwise_conv2.bias")
            norm(q + "norm_out", p + "norm_out")
        self._names = names
        return P

    def grad_buffers(self, P, device):
        """one zero-filled fp32 buffer holding every parameter gradient in PACKED layout (views by packed key)"""
        if self._layout is None:
            off, lay = 0, {}
            for k, key in self._names.items():
                shape = tuple(P[k].shape)
                n = 1
                for s in shape:
                    n *= s
                lay[k] = (off, n, shape)
                off += _al4(n)
            self._layout, self._total = lay, off
        flat = torch.zeros(self._total, dtype=torch.float32, device=device)
        return flat, {k: flat[o:o + n].view(shape) for k, (o, n, shape) in self._layout.items()}

    def to_param_grads(self, G) -> Dict[str, torch.Tensor]:
        """packed-layout gradients -> {parameter name: gradient in the reference's state_dict layout}"""
        m = self.m
        H, Dh, d, Cc, F3 = m.n_heads, m.head_dim, m.d_model, m.subsampling_conv_channels, m.subsampling.feat_sub
        params = dict(m.named_parameters())
        out = {}
        sc_used = m.self_conditioning and m.n_layers > 1
        for k, key in self._names.items():
            if k.startswith("dec_rep") and not sc_used:
                continue  # unused parameter: .grad stays None, as in the reference
            g = G[k]
            if k == "sub_out_w":
                g = g.view(d, F3, Cc).permute(0, 2, 1).reshape(d, Cc * F3)
            elif k.endswith("qkv_w"):
                g = g.view(3, H, Dh, d).permute(1, 2, 0, 3).reshape(3 * H * Dh, d)
            out[key] = g.reshape(params[key].shape)
        return out

    # ---- small host-side state ----------------------------------------------------------------------------
    def rope(self, N, device):
        m = self.m
        if not m.use_rotary:
            return None, None
        rp = m.rotary_pos_emb
        key = (N, str(device), rp.inv_freq._version, rp.rotary_interpolation_factor._version)
        if self._rope.get("key") != key:
            interp = float(rp.rotary_interpolation_factor)  # host sync, once per (N, buffer version)
            cos, sin = ops.rope_table(rp.inv_freq.to(device=device, dtype=torch.float32).contiguous(), interp, N)
            self._rope = {"key": key, "cos": cos, "sin": sin}
        return self._rope["cos"], self._rope["sin"]

    @staticmethod
    def brn_clamps(bn):
        """rmax, dmax of batchrenorm.py:41-50 from num_batches_tracked (host mirror: no sync per step)"""
        ver = bn.num_batches_tracked._version
        if getattr(bn, "_nbt_ver", None) != ver:
            bn._nbt_host = int(bn.num_batches_tracked)  # sync only when somebody else changed the buffer
        n = float(bn._nbt_host)
        rmax = min(max(2.0 / 35000.0 * n + 25.0 / 35.0, 1.0), 3.0)
        dmax = min(max(5.0 / 20000.0 * n - 25.0 / 20.0, 0.0), 5.0)
        return rmax, dmax

    @staticmethod
    def brn_tick(bn):
        bn.num_batches_tracked += 1
        bn._nbt_host += 1
        bn._nbt_ver = bn.num_batches_tracked._version

    # ---- forward -----------------------------------------------------------------------------------------
    def forward(self, spec: torch.Tensor, save: bool = True, tok_lens=None, brn_eval: bool = False, return_logits: bool = False):
        """spec [B,F,T] fp32 -> (log-probs [B,N,V1] fp32, argmax int32 [B,N], ctx for backward or None).
        tok_lens (host ints, valid tokens per recording, or None): the padded-batch path of sconformer_xl.py:204-215 —
        key-padding mask and zeroed rows in attention (attention.py:511,541), zeroed GLU output in the conv module
        (convolution.py:107-110); BatchRenorm statistics still cover every position, as in the reference.
        brn_eval: eval()-mode BatchRenorm (running statistics, batchrenorm.py:86-91; nothing is updated) with the graph kept
        for the backward (dynamic_eval.py).  return_logits: stop before the log-softmax (decoder.py:26-27)."""
        m = self.m
        dev = spec.device
        B, Fdim, Tn = spec.shape
        d, H, Dh, Cc, F3 = m.d_model, m.n_heads, m.head_dim, m.subsampling_conv_channels, m.subsampling.feat_sub
        V1 = m.decoder.num_classes
        kind = m.default_norm_name
        eps = 1e-8 if kind == "rms_norm" else 1e-5
        P = self.pack(dev)
        S: Dict = {"P": P, "spec": spec, "B": B, "T": Tn}
        lens_dev = None
        SILU, GELU = L.ACT_SILU, L.ACT_GELU_TANH

        def ln(x, pk, f32=False):
            o32, lo = ops.layernorm(x, P[pk + "_w"], P[pk + "_b"], eps=eps, kind=kind, out_f32=f32, lo_dtype=None if f32 else BF)
            return o32 if f32 else lo

        # subsampling.  Level 1 (conv0 + SiLU + depthwise): the fused kernel keeps the 160x-expanded conv0 activation in
        # shared memory and the fused backward recomputes it from the spectrogram, so it never exists in HBM.
        if Cc % 64 == 0:
            s1 = None
            d1 = ops.subsample_conv0_dw(spec, P["conv0_w"], P["conv0_b"], P["dw1_w"], P["dw1_b"])   # [B,T2,F2,C]
        else:
            s1 = ops.subsample_conv0(spec, P["conv0_w"], P["conv0_b"], out_dtype=BF)        # [B,T1,F1,C]
            d1 = ops.subsample_dwconv(s1, P["dw1_w"], P["dw1_b"])                            # [B,T2,F2,C]
        a1, p1 = T.gemm_act_pre(d1.view(-1, Cc), P["pw1_w"], P["pw1_b"], SILU)
        a1 = a1.view(d1.shape)
        d2 = ops.subsample_dwconv(a1, P["dw2_w"], P["dw2_b"])                                # [B,N,F3,C]
        N = d2.shape[1]
        M = B * N
        if tok_lens is not None and any(int(n) != N for n in tok_lens):
            lens_dev = torch.tensor([int(n) for n in tok_lens], dtype=torch.int32, device=dev)
        S["tok_lens"], S["lens_dev"] = ([int(n) for n in tok_lens] if lens_dev is not None else None), lens_dev
        a2, p2 = T.gemm_act_pre(d2.view(-1, Cc), P["pw2_w"], P["pw2_b"], SILU)
        x = ops.gemm(a2.view(M, F3 * Cc), P["sub_out_w"], out_dtype=torch.float32)           # [M,d] fp32
        S.update(s1=s1, d1=d1, p1=p1, a1=a1, d2=d2, p2=p2, a2=a2, N=N)
        cos, sin = self.rope(N, dev)
        S["cos"], S["sin"] = cos, sin

        layers: List[Dict] = []
        for l, layer in enumerate(m.layers):
            q = f"layers.{l}."
            R: Dict = {}
            for ff in ("ff1", "attn", "conv", "ff2"):
                if ff in ("ff1", "ff2"):
                    a = ln(x, q + ff + "_norm")
                    hact, hpre = T.gemm_act_pre(a, P[q + ff + "_fc1_w"], P[q + ff + "_fc1_b"], GELU)
                    xn = ops.gemm(hact, P[q + ff + "_fc2_w"], bias=P[q + ff + "_fc2_b"], resid=x, alpha=0.5)
                    R[ff] = dict(x=x, a=a, hpre=hpre, hact=hact)
                elif ff == "attn":
                    a = ln(x, q + "attn_norm")
                    if lens_dev is not None:
                        T.mask_rows_(a, lens_dev, B, N)
                    qkv = ops.gemm(a, P[q + "qkv_w"])
                    qq, kk, vv = ops.rope_split(qkv, B, N, H, Dh, cos, sin)
                    o, lse = T.attention_train(qq, kk, vv, lens_dev)
                    if lens_dev is not None:
                        T.mask_rows_(o, lens_dev, B, N)
                    xn = ops.gemm(o.view(M, d), P[q + "out_w"], resid=x, alpha=1.0)
                    R[ff] = dict(x=x, a=a, q=qq, k=kk, v=vv, o=o, lse=lse)
                else:
                    bn = layer.conv.fn.batch_norm
                    a = ln(x, q + "conv_norm")
                    u = ops.gemm(a, P[q + "pw1_w"], bias=P[q + "pw1_b"])
                    g = ops.glu(u) if lens_dev is None else T.glu_masked(u, lens_dev, B, N)
                    c, sums = T.dwconv1d_fwd(g.view(B, N, d), P[q + "dw_w"], P[q + "dw_b"], stats=not brn_eval)
                    if brn_eval:  # (c - running_mean) / running_std * weight + bias as the affine c*A + Bc; the backward reads
                        rm, rs = bn.running_mean.float(), bn.running_std.float()   # stats = mu, sigma, r = 1, d = 0, s = 0
                        A = (P[q + "brn_w"] / rs).contiguous()
                        Bc = (P[q + "brn_b"] - rm * A).contiguous()
                        stats = torch.stack([rm, rs, torch.ones_like(rm), torch.zeros_like(rm), torch.zeros_like(rm)]).contiguous()
                    else:
                        rmax, dmax = self.brn_clamps(bn)
                        A, Bc, stats = T.brn_train_stats(sums, M, bn.running_mean, bn.running_std, bn.eps, rmax, dmax, bn.momentum,
                                                         P[q + "brn_w"], P[q + "brn_b"])
                        self.brn_tick(bn)
                    y = T.affine_silu(c, A, Bc)
                    xn = ops.gemm(y.view(M, d), P[q + "pw2_w"], bias=P[q + "pw2_b"], resid=x, alpha=1.0)
                    R[ff] = dict(x=x, a=a, u=u, g=g, c=c, A=A, Bc=Bc, stats=stats, y=y, brn_eval=brn_eval)
                x = xn
            R["x_pre_out"] = x
            x = ln(x, q + "norm_out", f32=True)
            if l != m.n_layers - 1 and m.self_conditioning:  # sconformer_xl.py:241-243
                a6 = ln(x, "dec_norm") if m.decoder_norm else T.scale_cast(x, 1.0)
                logits = ops.gemm(a6, P["dec_ff_w"], bias=P["dec_ff_b"])
                prob = ops.softmax(logits)
                xn = ops.gemm(prob, P["dec_rep_w"], bias=P["dec_rep_b"], resid=x, alpha=1.0)
                R["sc"] = dict(x=x, a=a6, p=prob)
                x = xn
            layers.append(R)
        S["layers"] = layers
        if m.legasee_double_norm and m.decoder_norm:  # sconformer_xl.py:246
            S["x_dn"] = x
            x = ln(x, "dec_norm", f32=True)
        S["x_dec"] = x
        a = ln(x, "dec_norm") if m.decoder_norm else T.scale_cast(x, 1.0)
        S["a_dec"] = a
        lp = ops.gemm(a, P["dec_ff_w"], bias=P["dec_ff_b"], out_dtype=torch.float32)
        S["logits_out"] = return_logits
        if return_logits:
            return lp.view(B, N, V1), None, (S if save else None)
        am = ops.log_softmax_argmax_(lp)  # in place: logits -> log-probs
        S["lp"] = lp
        return lp.view(B, N, V1), am.view(B, N), (S if save else None)

    # ---- backward ----------------------------------------------------------------------------------------
    def backward(self, S: Dict, dlp: torch.Tensor) -> Dict[str, torch.Tensor]:
        m = self.m
        P = S["P"]
        dev = dlp.device
        B, N = S["B"], S["N"]
        M = B * N
        d, H, Dh, Cc, F3 = m.d_model, m.n_heads, m.head_dim, m.subsampling_conv_channels, m.subsampling.feat_sub
        V1 = m.decoder.num_classes
        kind = m.default_norm_name
        eps = 1e-8 if kind == "rms_norm" else 1e-5
        flat, G = self.grad_buffers(P, dev)
        SILU = L.ACT_SILU
        handles = []
        trace = self.trace  # debugging aid (tools/bwd_trace_check.py): list receiving (stage, clone) pairs, or None

        def ln_bwd(x, dy, pk, dx, accumulate, cast_scale=None):
            """cast_scale: also return bf16(cast_scale * dx), the dY operand of the sub-layer processed next"""
            return T.layernorm_bwd(x, dy, P[pk + "_w"], dx, G[pk + "_w"], G.get(pk + "_b"), eps=eps, kind=kind, accumulate=accumulate,
                                   cast_scale=cast_scale)

        def linear_bwd(dy, x_in, wk, bk=None, need_dx=True, **kw):
            """dy [M,out] bf16, x_in [M,in] bf16: weight (+bias) gradients, returns dx [M,in] bf16"""
            T.wgrad(dy, x_in, G[wk])
            if bk is not None and P.get(bk) is not None:
                T.colsum_(G[bk], dy)
            return T.dgrad(dy, P[wk], **kw) if need_dx else None

        dlp = dlp.contiguous().view(M, V1).to(torch.float32)
        dl = T.scale_cast(dlp, 1.0) if S.get("logits_out") else T.log_softmax_bwd(S["lp"], dlp)   # [M,V1] bf16
        if trace is not None:
            trace += [("dlp (CTC gradient)", dlp.clone()), ("dl (log-softmax backward, bf16)", dl.clone())]
        da = linear_bwd(dl, S["a_dec"], "dec_ff_w", "dec_ff_b")
        dx = torch.empty(M, d, dtype=torch.float32, device=dev)       # gradient of the residual stream
        if m.decoder_norm:
            ln_bwd(S["x_dec"], da, "dec_norm", dx, accumulate=False)
        else:
            dx.zero_()
            T.add_bf16_(dx, da)
        if "x_dn" in S:
            ln_bwd(S["x_dn"], dx, "dec_norm", dx, accumulate=False)   # in place: every row is read before it is written

        for l in range(m.n_layers - 1, -1, -1):
            q = f"layers.{l}."
            R = S["layers"][l]
            if "sc" in R:
                sc = R["sc"]
                dy = T.scale_cast(dx, 1.0)
                dp = linear_bwd(dy, sc["p"], "dec_rep_w", "dec_rep_b")
                dlg = T.softmax_bwd(sc["p"], dp, out=dp)
                da6 = linear_bwd(dlg, sc["a"], "dec_ff_w", "dec_ff_b")
                if m.decoder_norm:
                    ln_bwd(sc["x"], da6, "dec_norm", dx, accumulate=True)
                else:
                    T.add_bf16_(dx, da6)
            _, dy_next = ln_bwd(R["x_pre_out"], dx, q + "norm_out", dx, accumulate=False, cast_scale=0.5)  # ff2 comes next
            if trace is not None:
                trace += [(f"{q}dx after norm_out (fp32)", dx.clone()), (f"{q}dy into ff2 (bf16)", dy_next.clone())]
            nxt = {"ff2": 1.0, "conv": 1.0, "attn": 0.5}  # residual scale of the sub-layer that FOLLOWS in the backward order
            for ff in ("ff2", "conv", "attn", "ff1"):
                r = R[ff]
                cs = nxt.get(ff)  # None after ff1: the next consumer is the previous layer's self-conditioning / norm_out
                if ff in ("ff1", "ff2"):
                    dy = dy_next
                    dh = linear_bwd(dy, r["hact"], q + ff + "_fc2_w", q + ff + "_fc2_b", aux=r["hpre"], epi=L.EPI_GELU_BWD)
                    da = linear_bwd(dh, r["a"], q + ff + "_fc1_w", q + ff + "_fc1_b")
                    res = ln_bwd(r["x"], da, q + ff + "_norm", dx, accumulate=True, cast_scale=cs)
                    dy_next = res[1] if cs is not None else None
                elif ff == "conv":
                    dy = dy_next
                    dyy = linear_bwd(dy, r["y"].view(M, d), q + "pw2_w", q + "pw2_b")
                    dc = T.brn_silu_bwd(r["c"], dyy.view(B, N, d), r["A"], r["Bc"], r["stats"], P[q + "brn_w"], G[q + "brn_w"],
                                        G[q + "brn_b"], eval_mode=r["brn_eval"])
                    if trace is not None:
                        trace += [(f"{q}conv dy (bf16)", dy.clone()), (f"{q}conv dyy (bf16)", dyy.clone()), (f"{q}conv dc (bf16)", dc.clone())]
                    T.dwconv1d_bwd_weight_(r["g"].view(B, N, d), dc, G[q + "dw_w"], G[q + "dw_b"])
                    dg = T.dwconv1d_bwd_data(dc, P[q + "dw_w"])
                    if S["lens_dev"] is not None:
                        T.mask_rows_(dg, S["lens_dev"], B, N)
                    du = T.glu_bwd(r["u"], dg.view(M, d))
                    da = linear_bwd(du, r["a"], q + "pw1_w", q + "pw1_b")
                    _, dy_next = ln_bwd(r["x"], da, q + "conv_norm", dx, accumulate=True, cast_scale=cs)
                else:
                    dy = dy_next
                    do = linear_bwd(dy, r["o"].view(M, d), q + "out_w")
                    dq, dk, dv = T.attention_bwd(r["q"], r["k"], r["v"], r["o"].view(B, N, H, Dh), do.view(B, N, H, Dh), r["lse"],
                                                   lens=S["tok_lens"])
                    dqkv = T.rope_bwd_merge(dq, dk, dv, S["cos"], S["sin"])
                    da = linear_bwd(dqkv, r["a"], q + "qkv_w")
                    _, dy_next = ln_bwd(r["x"], da, q + "attn_norm", dx, accumulate=True, cast_scale=cs)
            S["layers"][l] = None  # release this layer's activations
            if self.dp_group is not None and self.dp_overlap:  # this layer's parameter gradients are final: reduce them behind
                handles.append(self._reduce_slice(flat, q))                # the rest of the backward

        # subsampling
        dy = T.scale_cast(dx, 1.0)
        dp2 = linear_bwd(dy, S["a2"].view(M, F3 * Cc), "sub_out_w", aux=S["p2"].view(M, F3 * Cc), epi=L.EPI_SILU_BWD).view(-1, Cc)
        dd2 = linear_bwd(dp2, S["d2"].view(-1, Cc), "pw2_w", "pw2_b").view(S["d2"].shape)
        T.subsample_dwconv_bwd_weight_(S["a1"], dd2, G["dw2_w"], G["dw2_b"])
        da1 = T.subsample_dwconv_bwd_data(dd2, P["dw2_w"], S["a1"].shape[1], S["a1"].shape[2])
        dp1 = T.act_bwd(S["p1"], da1.view(-1, Cc), SILU)
        dd1 = linear_bwd(dp1, S["d1"].view(-1, Cc), "pw1_w", "pw1_b").view(S["d1"].shape)
        if S["s1"] is None:
            T.subsample_l1_bwd_(S["spec"], P["conv0_w"], P["conv0_b"], P["dw1_w"], dd1, G["conv0_w"], G["conv0_b"], G["dw1_w"],
                                G["dw1_b"])
        else:
            T.subsample_dwconv_bwd_weight_(S["s1"], dd1, G["dw1_w"], G["dw1_b"])
            ds1 = T.subsample_dwconv_bwd_data(dd1, P["dw1_w"], S["s1"].shape[1], S["s1"].shape[2])
            T.subsample_conv0_bwd_(S["spec"], P["conv0_w"], P["conv0_b"], ds1, G["conv0_w"], G["conv0_b"])

        if self.dp_group is not None and not self.dp_overlap:
            self.reduce_gradients(flat)
        elif self.dp_group is not None:
            handles.append(self._reduce_slice(flat, None))  # subsampling + decoder (accumulated across all layers)
            for h in handles:
                h.wait()
            if self.dp_average:
                import torch.distributed as dist
                flat.mul_(1.0 / dist.get_world_size(self.dp_group))
        self.last_flat_grad = flat
        return self.to_param_grads(G)

    def _reduce_slice(self, flat: torch.Tensor, layer_prefix):
        """asynchronous all-reduce (sum) of the contiguous slice of the flat gradient buffer that belongs to one layer
        (`layers.<l>.`) or, with None, to everything outside the layer stack.  NCCL orders it after the kernels already
        enqueued on the current stream and runs it on its own stream, overlapping the remaining backward."""
        import torch.distributed as dist
        keys = [k for k in self._layout if (k.startswith(layer_prefix) if layer_prefix else not k.startswith("layers."))]
        lo = min(self._layout[k][0] for k in keys)
        hi = max(self._layout[k][0] + _al4(self._layout[k][1]) for k in keys)
        return dist.all_reduce(flat[lo:hi], group=self.dp_group, async_op=True)

    def reduce_gradients(self, flat: torch.Tensor) -> None:
        """data-parallel training: ONE all-reduce over the flat packed gradient buffer (NCCL over NVLink on the GPU
        box; sum, then divided by the world size like DistributedDataParallel when dp_average)."""
        if self.dp_group is None:
            return
        import torch.distributed as dist
        dist.all_reduce(flat, group=self.dp_group)
        if self.dp_average:
            flat.mul_(1.0 / dist.get_world_size(self.dp_group))


class _EncoderFn(torch.autograd.Function):
    """log_probs = f(spec; parameters): forward / backward are TrainEngine.forward / backward."""

    @staticmethod
    def forward(ctx, engine: TrainEngine, spec: torch.Tensor, tok_lens, brn_eval, return_logits, *params):
        with torch.no_grad():
            lp, am, S = engine.forward(spec, save=True, tok_lens=tok_lens, brn_eval=brn_eval, return_logits=return_logits)
        ctx.engine, ctx.S = engine, S
        ctx.names = [n for n, _ in engine.m.named_parameters()]
        engine.m.last_argmax = am
        return lp

    @staticmethod
    def backward(ctx, dlp):
        S, ctx.S = ctx.S, None
        if S is None:
            raise RuntimeError("lcasr_b200: backward through the encoder a second time (activations were freed)")
        with torch.no_grad():
            grads = ctx.engine.backward(S, dlp)
        return (None, None, None, None, None) + tuple(grads.get(n) for n in ctx.names)


def train_forward(model, audio_signal: torch.Tensor, tok_lens=None, brn_eval: bool = False, return_logits: bool = False) -> torch.Tensor:
    if getattr(model, "_train_engine", None) is None:
        model._train_engine = TrainEngine(model)
    eng = model._train_engine
    params = [p for _, p in model.named_parameters()]
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _EncoderFn.apply(eng, audio_signal, tok_lens, brn_eval, return_logits, *params)
    lp, am, _ = eng.forward(audio_signal, save=False, tok_lens=tok_lens, brn_eval=brn_eval, return_logits=return_logits)
    model.last_argmax = am
    return lp
