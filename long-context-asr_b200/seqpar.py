"""Sequence-parallel forward of SCConformerXL: one recording, its N tokens split into P
contiguous blocks, one block per rank (one process per GPU).

What crosses ranks (everything else is token-local):
  * attention  — K and V of every rank are all-gathered (NCCL over NVLink / NVSwitch) and each rank
                 attends its own query block against the whole recording (`lcasr_attention_cross`);
                 rotary tables are built with the block's position offset;
  * conv module — the depthwise k=9 convolution needs (k-1)/2 = 4 tokens of the post-GLU tensor from each
                 neighbour (zeros at the true sequence ends);
  * greedy decode — per-frame argmax ids are gathered so the collapse sees the seams.
The 8x subsampling needs NO communication: every rank reads its slice of the spectrogram plus 8 input
frames (= 1 token) of left context; the zero padding of each conv level only contaminates output row 0,
which is that context token and is dropped (frame counts are multiples of 8, so the right edge of a slice
never touches padding: out = floor((L-1)/2)+1 with L even reads inputs 2o-1..2o+1 <= L-1).

The same code runs with `DistComm` (torch.distributed, real ranks) and `LocalComm` (all P blocks in one
process on one GPU, used by the single-GPU parity test: results are bit-identical to the P=1 forward
because every kernel is row-independent and K/V tiles are visited in the same order).
PyTorch here is plumbing (memory, NCCL); every FLOP runs in liblcasr_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops


def split_tokens(n_tokens: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) token blocks, sizes differing by at most one (larger blocks first)."""
    base, rem = divmod(n_tokens, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((s, s + n))
        s += n
    return out


def frame_slice(start_tok: int, end_tok: int) -> Tuple[int, int, int]:
    """Input frames a rank needs for tokens [start, end): one token (8 frames) of left context unless it owns
    token 0.  Returns (first_frame, last_frame_exclusive, tokens_to_drop_at_the_front)."""
    drop = 1 if start_tok > 0 else 0
    return 8 * (start_tok - drop), 8 * end_tok, drop


class LocalComm:
    """All ranks live in this process: 'collectives' are list operations (single-GPU emulation)."""

    def __init__(self, world: int):
        self.world = world
        self.local_ranks = list(range(world))

    def all_gather_cat(self, xs: Sequence[torch.Tensor], sizes=None) -> List[torch.Tensor]:
        cat = torch.cat(list(xs), dim=0)
        return [cat for _ in xs]

    def halo_exchange(self, firsts: Sequence[torch.Tensor], lasts: Sequence[torch.Tensor]):
        """firsts[r]/lasts[r]: first/last h rows of rank r's block.  Returns per rank (left_halo, right_halo):
        the previous rank's last rows / next rank's first rows, zeros at the sequence ends."""
        out = []
        for r in range(self.world):
            left = lasts[r - 1] if r > 0 else torch.zeros_like(firsts[r])
            right = firsts[r + 1] if r + 1 < self.world else torch.zeros_like(lasts[r])
            out.append((left, right))
        return out


class DistComm:
    """One rank per process over torch.distributed (backend nccl on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local_ranks = [self.rank]

    def all_gather_cat(self, xs: Sequence[torch.Tensor], sizes=None) -> List[torch.Tensor]:
        """concatenation of every rank's rows.  `sizes` (rows per rank) is known to the caller from the static token
        split; passing it avoids a size exchange and — more importantly — the host synchronisation that reading it
        back costs on EVERY collective (3 per layer: the host could never run ahead of the GPU)."""
        (x,) = xs
        if sizes is None:
            n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
            szt = [torch.zeros_like(n) for _ in range(self.world)]
            self.dist.all_gather(szt, n, group=self.group)
            sizes = [int(s.item()) for s in szt]
        sizes = [int(v) for v in sizes]
        if len(set(sizes)) == 1:
            out = torch.empty((sizes[0] * self.world,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            self.dist.all_gather_into_tensor(out, x.contiguous(), group=self.group)
            return [out]
        mx = max(sizes)  # ragged blocks: pad to the largest, gather, trim
        pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[: x.shape[0]] = x
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(bufs, pad, group=self.group)
        return [torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)]

    def halo_exchange(self, firsts, lasts):
        (first,), (last,) = firsts, lasts
        both = torch.stack([first, last])  # [2, h, d]
        bufs = [torch.empty_like(both) for _ in range(self.world)]
        self.dist.all_gather(bufs, both.contiguous(), group=self.group)
        left = bufs[self.rank - 1][1] if self.rank > 0 else torch.zeros_like(first)
        right = bufs[self.rank + 1][0] if self.rank + 1 < self.world else torch.zeros_like(last)
        return [(left, right)]


@torch.no_grad()
def forward_sequence_parallel(model, spec: torch.Tensor, comm, return_logits: bool = False):
    """spec [1, feat_in, T] fp32 CUDA (the whole recording; each rank only reads its slice), T % 8 == 0.
    Returns per LOCAL rank: (log-probs [n_r, V1] fp32, argmax int32 [n_r], (start_tok, end_tok)); the gathered
    argmax ids of the whole recording (for the greedy collapse) are returned as the last element."""
    assert spec.dim() == 3 and spec.shape[0] == 1, "sequence parallelism shards ONE recording"
    T = spec.shape[-1]
    assert T % 8 == 0, "sequence-parallel subsampling assumes a frame count that is a multiple of 8"
    device = spec.device
    model._ensure_built(device)
    P = model._packed
    cdt = model.compute_dtype
    d, H, Dh, Cc = model.d_model, model.n_heads, model.head_dim, model.subsampling_conv_channels
    V1, F3, KS = model.decoder.num_classes, model.subsampling.feat_sub, model.conv_kernel_size
    rms = model.default_norm_name == "rms_norm"
    eps, kind = (1e-8, "rms_norm") if rms else (1e-5, "layer_norm")
    N = ops.out_length(T)
    blocks = split_tokens(N, comm.world)
    sizes = [e - b for b, e in blocks]  # rows every rank contributes to a gather: static, no size exchange
    gi, ai = model._impl
    halo = (KS - 1) // 2

    def norm(x, wkey, out_f32=False, lo=True):
        return ops.layernorm(x, P[wkey + "_w"], P.get(wkey + "_b"), eps, kind, out_f32=out_f32, lo_dtype=cdt if lo else None)

    # ---- subsampling: slice + one token of left context, no communication ----
    xs, offs = [], []
    for r in comm.local_ranks:
        s_tok, e_tok = blocks[r]
        f0, f1, drop = frame_slice(s_tok, e_tok)
        sl = spec[:, :, f0:f1].contiguous()
        if cdt == torch.bfloat16 and Cc % 64 == 0:
            a = ops.subsample_conv0_dw(sl, P["conv0_w"], P["conv0_b"], P["dw1_w"], P["dw1_b"])
        else:
            a = ops.subsample_conv0(sl, P["conv0_w"], P["conv0_b"], cdt)
            a = ops.subsample_dwconv(a, P["dw1_w"], P["dw1_b"])
        B_, T2, F2, _ = a.shape
        a = ops.gemm(a.view(-1, Cc), P["pw1_w"], bias=P["pw1_b"], act=L.ACT_SILU, impl=gi).view(B_, T2, F2, Cc)
        a = ops.subsample_dwconv(a, P["dw2_w"], P["dw2_b"])
        B_, T3, F3_, _ = a.shape
        a = ops.gemm(a.view(-1, Cc), P["pw2_w"], bias=P["pw2_b"], act=L.ACT_SILU, impl=gi)
        x = ops.gemm(a.view(T3, F3_ * Cc), P["sub_out_w"], out_dtype=torch.float32, impl=gi)
        x = x[drop:].contiguous()
        assert x.shape[0] == e_tok - s_tok
        xs.append(x)
        offs.append(s_tok)

    interp = 1.0
    if model.use_rotary:  # the buffer lives on the GPU: read it back once per buffer version, not once per forward
        buf = model.rotary_pos_emb.rotary_interpolation_factor
        if getattr(model, "_interp_cache", (None, None))[0] != (buf.data_ptr(), buf._version):
            model._interp_cache = ((buf.data_ptr(), buf._version), float(buf))
        interp = model._interp_cache[1]
    tables = [ops.rope_table(P["inv_freq"], interp, x.shape[0], offset=o) if model.use_rotary else (None, None)
              for x, o in zip(xs, offs)]

    def ffn(x, p, name):
        _, a = norm(x, p + name + "_norm")
        hmid = ops.gemm(a, P[p + name + "_fc1_w"], bias=P.get(p + name + "_fc1_b"), act=L.ACT_GELU_TANH, impl=gi)
        ops.gemm(hmid, P[p + name + "_fc2_w"], bias=P.get(p + name + "_fc2_b"), resid=x, alpha=0.5, impl=gi, out=x)

    for l in range(model.n_layers):
        p = f"layers.{l}."
        for x in xs:
            ffn(x, p, "ff1")
        # attention: local queries against the all-gathered keys / values
        qs, ks, vs = [], [], []
        for x, (cos, sin) in zip(xs, tables):
            _, a = norm(x, p + "attn_norm")
            qkv = ops.gemm(a, P[p + "qkv_w"], impl=gi)
            q, k, v = ops.rope_split(qkv, 1, x.shape[0], H, Dh, cos, sin)
            qs.append(q); ks.append(k.view(-1, d)); vs.append(v.view(-1, d))
        kf = comm.all_gather_cat(ks, sizes)
        vf = comm.all_gather_cat(vs, sizes)
        for x, q, kk, vv in zip(xs, qs, kf, vf):
            o = ops.attention_cross(q, kk.view(1, -1, H, Dh), vv.view(1, -1, H, Dh), impl=ai)
            ops.gemm(o.view(-1, d), P[p + "out_w"], resid=x, alpha=1.0, impl=gi, out=x)
        # convolution module with a (k-1)/2-token halo of the post-GLU tensor
        gs = []
        for x in xs:
            _, a = norm(x, p + "conv_norm")
            gs.append(ops.glu(ops.gemm(a, P[p + "pw1_w"], bias=P[p + "pw1_b"], impl=gi)))
        halos = comm.halo_exchange([g[:halo] for g in gs], [g[-halo:] for g in gs])
        for x, g, (left, right) in zip(xs, gs, halos):
            ext = torch.cat([left, g, right], dim=0)
            c = ops.dwconv_brn_silu(ext.view(1, -1, d), P[p + "dw_w"], P[p + "dw_b"], P[p + "brn_mean"], P[p + "brn_std"],
                                    P[p + "brn_w"], P[p + "brn_b"]).view(-1, d)
            ops.gemm(c[halo: halo + g.shape[0]], P[p + "pw2_w"], bias=P[p + "pw2_b"], resid=x, alpha=1.0, impl=gi, out=x)
        for i, x in enumerate(xs):
            ffn(x, p, "ff2")
            xs[i], _ = norm(x, p + "norm_out", out_f32=True, lo=False)
        if l != model.n_layers - 1 and model.self_conditioning:
            for x in xs:
                a = norm(x, "dec_norm")[1] if model.decoder_norm else x.to(cdt)
                pr = ops.softmax(ops.gemm(a, P["dec_ff_w"], bias=P["dec_ff_b"], impl=gi))
                ops.gemm(pr, P["dec_rep_w"], bias=P["dec_rep_b"], resid=x, alpha=1.0, impl=gi, out=x)

    outs, ams = [], []
    for x in xs:
        if model.legasee_double_norm and model.decoder_norm:
            x, _ = norm(x, "dec_norm", out_f32=True, lo=False)
        a = norm(x, "dec_norm")[1] if model.decoder_norm else x.to(cdt)
        logits = ops.gemm(a, P["dec_ff_w"], bias=P["dec_ff_b"], out_dtype=torch.float32, impl=gi)
        am = None if return_logits else ops.log_softmax_argmax_(logits)
        outs.append(logits); ams.append(am)
    am_full = None if return_logits else comm.all_gather_cat(ams, sizes)[0]
    return [(o, a, blocks[r]) for o, a, r in zip(outs, ams, comm.local_ranks)], am_full


# ------------------------------------------------------------------------------------------------------
# Native driver (csrc/seqpar.cu): the whole per-rank forward is ONE C-ABI call; K/V blocks travel as NCCL
# point-to-point transfers in ring order on a side stream while attention runs on the blocks already present
# (exact merge of per-block partial results), neighbour-only halo exchange, argmax ids exchanged at the end.
# The Python driver above stays as the readable statement of the algorithm (and runs under gloo on CPU-only hosts).
# ------------------------------------------------------------------------------------------------------
import ctypes as _C
import os as _os


def _nccl_path() -> bytes:
    cand = _os.path.join(_os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2")
    return _os.path.abspath(cand).encode() if _os.path.exists(cand) else b""


class NativeComm:
    """NCCL communicator owned by liblcasr_b200.so, bootstrapped over an existing torch.distributed group
    (rank 0 creates the NCCL unique id, a broadcast distributes it; one process per GPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = _C.create_string_buffer(128)
        if self.rank == 0:
            L.call("lcasr_comm_unique_id", buf, _nccl_path())
        backend = dist.get_backend(group)
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        t = t.to(dev) if backend == "nccl" else t
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idb = bytes(t.cpu().numpy().tobytes())
        h = L.vp()
        torch.cuda.synchronize()
        L.call("lcasr_comm_create", idb, self.rank, self.world, _nccl_path(), _C.byref(h))
        self._h = h.value

    def close(self):
        if getattr(self, "_h", None):
            L.lib.lcasr_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _sp_buffers(model, key, nbytes, device):
    cache = model.__dict__.setdefault("_sp_ws", {})
    ws = cache.get(key)
    if ws is None or ws.numel() < nbytes or ws.device != device:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        cache[key] = ws
    return ws


@torch.no_grad()
def forward_sequence_parallel_native(model, spec: torch.Tensor, comm: NativeComm, return_logits: bool = False):
    """This rank's share of SCConformerXL.forward for ONE recording.  spec [1, feat_in, T] fp32 CUDA (the whole recording
    on every rank; only the rank's slice is read), T % 8 == 0.  Returns (log-probs [n_rank, V1] fp32, argmax ids of the
    WHOLE recording [N] int32 (None with return_logits), (start_tok, end_tok))."""
    assert spec.dim() == 3 and spec.shape[0] == 1 and spec.is_cuda and spec.dtype == torch.float32 and spec.is_contiguous()
    T = spec.shape[-1]
    device = spec.device
    model._ensure_built(device)
    s, n = L.i64(), L.i64()
    L.call("lcasr_model_seqpar_block", model._handle, comm.world, comm.rank, T, _C.byref(s), _C.byref(n))
    V1 = model.decoder.num_classes
    out = torch.empty(n.value, V1, dtype=torch.float32, device=device)
    am = None if return_logits else torch.empty(T // 8, dtype=torch.int32, device=device)
    nbytes = int(L.lib.lcasr_model_seqpar_workspace_bytes(model._handle, comm.world, comm.rank, T))
    ws = _sp_buffers(model, ("native", comm.world, comm.rank, T), nbytes, device)
    L.call("lcasr_model_forward_seqpar", model._handle, comm._h, L.ptr(spec), T, L.ptr(out), L.ptr(am), int(return_logits),
           L.ptr(ws), ws.numel(), L.current_stream())
    return out, am, (s.value, s.value + n.value)


@torch.no_grad()
def forward_sequence_parallel_emulated(model, spec: torch.Tensor, world: int, return_logits: bool = False):
    """All `world` ranks of the native driver in this process on one GPU (copies instead of NCCL): (log-probs [N, V1],
    argmax [N])."""
    assert spec.dim() == 3 and spec.shape[0] == 1 and spec.is_cuda and spec.dtype == torch.float32 and spec.is_contiguous()
    T = spec.shape[-1]
    device = spec.device
    model._ensure_built(device)
    V1 = model.decoder.num_classes
    out = torch.empty(T // 8, V1, dtype=torch.float32, device=device)
    am = None if return_logits else torch.empty(T // 8, dtype=torch.int32, device=device)
    nbytes = int(L.lib.lcasr_model_seqpar_emulated_workspace_bytes(model._handle, world, T))
    if nbytes < 0:
        raise L.LcasrError(f"sequence-parallel plan rejected: {L.lib.lcasr_last_error().decode()}")
    ws = _sp_buffers(model, ("emulated", world, T), nbytes, device)
    L.call("lcasr_model_forward_seqpar_emulated", model._handle, world, L.ptr(spec), T, L.ptr(out), L.ptr(am), int(return_logits),
           L.ptr(ws), ws.numel(), L.current_stream())
    return out, am
