"""Long-form (moving-window) inference — drop-in for ``lcasr/eval/utils.py:45-111`` (``fetch_logits``), the caller of
the hot path in every published evaluation of the reference (eval/run.py:84).

The reference walks the recording window by window (batch 1), copies each ``[n, 4096]`` posterior block to the host,
and accumulates ``exp`` / counts there.  Here the windows are gathered into batches, run through ``SCConformerXL`` in a
few large forward calls, and ONE kernel (``lcasr_window_merge``) averages the probabilities of overlapping windows,
takes the logarithm and the per-frame argmax on the device.  ``fetch_logits`` returns what the reference returns (a
numpy ``[N, V+1]`` array of log-probabilities); ``transcribe_longform`` keeps everything on the GPU and returns token ids.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib as L
from . import ops


def plan_windows(spec_n: int, seq_len: int, overlap: int, downsampling_factor: int = 8) -> Tuple[List[Tuple[int, int]], int, int]:
    """(start, length) of every window the reference's loop processes (utils.py:66-83: it stops after the first
    window that is shorter than its predecessor), plus the effective seq_len / overlap."""
    if seq_len > spec_n:
        seq_len, overlap = spec_n, 0
    assert overlap / downsampling_factor == overlap // downsampling_factor, "Overlap must be a multiple of the downsampling factor"
    assert seq_len - overlap > 0, "overlap must be smaller than seq_len"
    wins, last_ulen, kill_next = [], None, False
    for i in range(0, spec_n, seq_len - overlap):
        u_len = min(seq_len, spec_n - i)
        if kill_next:
            break
        if last_ulen is not None and u_len < last_ulen:
            kill_next = True
        last_ulen = u_len
        wins.append((i, u_len))
    return wins, seq_len, overlap


def window_positions(wins, ds_lens, overlap: int) -> List[int]:
    """merged-frame position of every window, exactly like the running `logit_position` of utils.py:91-98"""
    pos, position = [], 0
    for k, ((i, u_len), ds_len) in enumerate(zip(wins, ds_lens)):
        ratio = u_len / ds_len
        overlap_ds = int(overlap / ratio)
        if i != 0:
            position -= overlap_ds
        pos.append(position)
        position += ds_len
    return pos


@torch.no_grad()
def _merged(model, spec: torch.Tensor, seq_len: int, overlap: int, want_logits: bool, max_batch: int):
    if spec.dim() == 2:
        spec = spec.unsqueeze(0)
    assert spec.dim() == 3 and spec.shape[0] == 1, "fetch_logits takes one recording [1, feat, T] (utils.py:46)"
    dev = spec.device if spec.is_cuda else torch.device("cuda", torch.cuda.current_device())
    spec = spec.to(dev, torch.float32)
    spec_n = spec.shape[-1]
    wins, seq_len, overlap = plan_windows(spec_n, seq_len, overlap, model.subsampling.subsampling_factor)
    ds_lens = [ops.out_length(u) for _, u in wins]
    pos = window_positions(wins, ds_lens, overlap)
    V1 = model.decoder.num_classes
    was_training = model.training
    model.eval()
    blocks = []
    full = [k for k, (_, u) in enumerate(wins) if u == seq_len]
    starts = torch.tensor([wins[k][0] for k in full], device=dev)
    for b0 in range(0, len(full), max_batch):  # equal-length windows: one batched forward per group
        idx = starts[b0:b0 + max_batch, None] + torch.arange(seq_len, device=dev)[None, :]
        batch = spec[0][:, idx].permute(1, 0, 2).contiguous()  # [k, feat, seq_len]
        blocks.append(model(batch)["final_posteriors"].reshape(-1, V1))
    for k, (i, u) in enumerate(wins):  # the (at most one) shorter tail window
        if u != seq_len:
            blocks.append(model(spec[:, :, i:i + u].contiguous())["final_posteriors"].reshape(-1, V1))
    model.train(was_training)
    logp = blocks[0] if len(blocks) == 1 else torch.cat(blocks, 0)
    row0, r = [], 0
    order = full + [k for k in range(len(wins)) if k not in set(full)]
    offs = {}
    for k in order:
        offs[k] = r
        r += ds_lens[k]
    row0 = torch.tensor([offs[k] for k in range(len(wins))], dtype=torch.int64, device=dev)
    wlen = torch.tensor(ds_lens, dtype=torch.int32, device=dev)
    wpos = torch.tensor(pos, dtype=torch.int32, device=dev)
    n_total = max(p + n for p, n in zip(pos, ds_lens))
    out = torch.empty(n_total, V1, dtype=torch.float32, device=dev) if want_logits else None
    am = torch.empty(n_total, dtype=torch.int32, device=dev)
    L.call("lcasr_window_merge", L.ptr(logp), V1, len(wins), L.ptr(row0), L.ptr(wlen), L.ptr(wpos), max(ds_lens), n_total,
           L.ptr(out), L.ptr(am), L.current_stream())
    return out, am


def plan_buffers(spec_n: int, seq_len: int, overlap: int) -> Tuple[List[Tuple[int, int, int, int]], int, int]:
    """(buffer_start, buffer_end, chunk_start, chunk_end) of every step of the reference's buffered mode
    (buffered_transcription.py:42-72): chunks of seq_len - overlap frames, each inside a seq_len-frame buffer reaching
    overlap/2 frames to both sides and shifted inwards at the edges; plus the effective seq_len / overlap."""
    if seq_len > spec_n:
        seq_len, overlap = spec_n, 0
    chunk = seq_len - overlap
    assert chunk > 0, "overlap must be smaller than seq_len"
    steps, c0, c1 = [], 0, chunk
    while True:
        s0, s1 = c0 - overlap // 2, c1 + overlap // 2
        if s0 < 0:
            s0, s1 = 0, seq_len
        elif s1 > spec_n:
            s0, s1 = spec_n - seq_len, spec_n
        steps.append((s0, s1, c0, c1))
        c0, c1 = c0 + chunk, min(c1 + chunk, spec_n)
        if c0 >= spec_n:
            return steps, seq_len, overlap


@torch.no_grad()
def _buffered(model, spec: torch.Tensor, seq_len: int, overlap: int, want_logits: bool, max_batch: int):
    if spec.dim() == 2:
        spec = spec.unsqueeze(0)
    assert spec.dim() == 3 and spec.shape[0] == 1, "fetch_logits takes one recording [1, feat, T]"
    dev = spec.device if spec.is_cuda else torch.device("cuda", torch.cuda.current_device())
    spec = spec.to(dev, torch.float32)
    spec_n = spec.shape[-1]
    steps, seq_len, overlap = plan_buffers(spec_n, seq_len, overlap)
    ds = model.subsampling.subsampling_factor
    assert overlap / ds == overlap // ds, "Overlap must be a multiple of the downsampling factor"
    V1 = model.decoder.num_classes
    n_rows = ops.out_length(seq_len)  # every buffer is seq_len frames long
    ratio = seq_len / n_rows
    was_training = model.training
    model.eval()
    blocks = []
    starts = torch.tensor([s[0] for s in steps], device=dev)
    for b0 in range(0, len(steps), max_batch):
        idx = starts[b0:b0 + max_batch, None] + torch.arange(seq_len, device=dev)[None, :]
        batch = spec[0][:, idx].permute(1, 0, 2).contiguous()  # [k, feat, seq_len]
        blocks.append(model(batch)["final_posteriors"].reshape(-1, V1))
    model.train(was_training)
    logp = blocks[0] if len(blocks) == 1 else torch.cat(blocks, 0)
    row0, wlen, wpos, position = [], [], [], 0
    for k, (s0, s1, c0, c1) in enumerate(steps):  # rows of the central chunk (buffered_transcription.py:84-90)
        r0, r1 = int((c0 - s0) / ratio), int((c1 - s0) / ratio)
        row0.append(k * n_rows + r0)
        wlen.append(r1 - r0)
        wpos.append(position)
        position += r1 - r0
    keep = [k for k in range(len(steps)) if wlen[k] > 0]
    assert keep, "recording too short for one output frame"
    row0_t = torch.tensor([row0[k] for k in keep], dtype=torch.int64, device=dev)
    wlen_t = torch.tensor([wlen[k] for k in keep], dtype=torch.int32, device=dev)
    wpos_t = torch.tensor([wpos[k] for k in keep], dtype=torch.int32, device=dev)
    out = torch.empty(position, V1, dtype=torch.float32, device=dev) if want_logits else None
    am = torch.empty(position, dtype=torch.int32, device=dev)
    L.call("lcasr_window_concat", L.ptr(logp), V1, len(keep), L.ptr(row0_t), L.ptr(wlen_t), L.ptr(wpos_t), max(wlen), position,
           L.ptr(out), L.ptr(am), L.current_stream())
    return out, am


def fetch_logits_buffered(args, model, spec: torch.Tensor, seq_len: int, overlap: int, tokenizer=None, use_tqdm: bool = True,
                          max_batch: int = 16):
    """Drop-in for ``lcasr/eval/buffered_transcription.py:11-97`` (``fetch_logits`` of the buffered mode): same arguments
    and result (numpy float32 [N, V+1]: the central-chunk log-probabilities of successive buffers, concatenated).  All
    buffers run through the encoder in batches of `max_batch`; one kernel gathers the kept rows and their argmax."""
    if seq_len == -1:
        seq_len = args.config["audio_chunking"]["size"]
    if overlap == -1 and seq_len <= spec.shape[-1]:
        overlap = args.config["audio_chunking"]["overlap"]
    out, _ = _buffered(model, spec, seq_len, overlap, True, max_batch)
    return out.cpu().numpy()


def transcribe_buffered(model, spec: torch.Tensor, seq_len: int, overlap: int, blank_id: Optional[int] = None,
                        max_batch: int = 16) -> List[int]:
    """buffered fetch_logits + GreedyCTCDecoder with only token ids leaving the GPU."""
    _, am = _buffered(model, spec, seq_len, overlap, False, max_batch)
    blank = model.decoder.num_classes - 1 if blank_id is None else blank_id
    toks, n = ops.greedy_collapse(am.view(1, -1), blank)
    return toks[0, : int(n[0])].tolist()


def fetch_logits(args, model, spec: torch.Tensor, seq_len: int, overlap: int, tokenizer=None, use_tqdm: bool = True,
                 max_batch: int = 16):
    """Same arguments and result as the reference (numpy float32 [N, V+1] log-probabilities of the merged windows);
    `seq_len` / `overlap` == -1 fall back to args.config['audio_chunking'] like utils.py:49,55."""
    if seq_len == -1:
        seq_len = args.config["audio_chunking"]["size"]
    if overlap == -1 and seq_len <= spec.shape[-1]:
        overlap = args.config["audio_chunking"]["overlap"]
    out, _ = _merged(model, spec, seq_len, overlap, True, max_batch)
    return out.cpu().numpy()


def transcribe_longform(model, spec: torch.Tensor, seq_len: int, overlap: int, blank_id: Optional[int] = None,
                        max_batch: int = 16) -> List[int]:
    """fetch_logits + GreedyCTCDecoder (eval/run.py:84-89) without the posteriors ever leaving the GPU: returns the
    greedy CTC token ids of the whole recording."""
    _, am = _merged(model, spec, seq_len, overlap, False, max_batch)
    blank = model.decoder.num_classes - 1 if blank_id is None else blank_id
    toks, n = ops.greedy_collapse(am.view(1, -1), blank)
    return toks[0, : int(n[0])].tolist()
