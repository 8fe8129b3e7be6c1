#!/usr/bin/env python
"""Run ONE hot-path operator a few times on cuda:0 (for ncu captures and quick timing).
  python tools/run_op.py attn --N 16384 --H 24 --Dh 32
  python tools/run_op.py gemm --M 16384 --Nout 3072 --K 768 [--act gelu] [--resid] [--out f32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lcasr_b200  # noqa: E402
from lcasr_b200 import ops, _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("op", choices=["attn", "gemm"])
    ap.add_argument("--N", type=int, default=16384)
    ap.add_argument("--H", type=int, default=24)
    ap.add_argument("--Dh", type=int, default=32)
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--M", type=int, default=16384)
    ap.add_argument("--Nout", type=int, default=3072)
    ap.add_argument("--K", type=int, default=768)
    ap.add_argument("--act", default="none")
    ap.add_argument("--resid", action="store_true")
    ap.add_argument("--out", default="bf16")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--check", action="store_true", help="attn: error against chunked fp32 softmax(QK^T)V")
    ap.add_argument("--qscale", type=float, default=1.0, help="attn: scale of q (larger = more peaked rows)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    if a.op == "attn":
        q, k, v = (torch.randn(a.B, a.N, a.H, a.Dh, generator=g).bfloat16().to(dev) for _ in range(3))
        q = (q.float() * a.qscale).bfloat16()
        if a.check:
            o = ops.attention(q, k, v, impl=L.ATTN_TCGEN05).float().reshape(a.B, a.N, a.H, a.Dh)
            kf, vf = k.float().transpose(1, 2), v.float().transpose(1, 2)
            worst, bias, sq, cnt = 0.0, 0.0, 0.0, 0
            for i in range(0, a.N, 2048):
                qf = q[:, i:i + 2048].float().transpose(1, 2)
                ref = torch.softmax(qf @ kf.transpose(-1, -2) / a.Dh ** 0.5, dim=-1) @ vf
                d = o[:, i:i + 2048].transpose(1, 2) - ref
                worst = max(worst, d.abs().max().item())
                bias += (d * ref.sign()).sum().item()
                sq += (d * d).sum().item()
                cnt += d.numel()
            print(f"attn check qscale {a.qscale}: max-abs {worst:.3e}  rms {(sq / cnt) ** 0.5:.3e}  mean signed (towards |ref|) {bias / cnt:.3e}")
        fn = lambda: ops.attention(q, k, v, impl=L.ATTN_TCGEN05)
        flops = 4.0 * a.B * a.H * a.N * a.N * a.Dh
    else:
        x = torch.randn(a.M, a.K, generator=g).bfloat16().to(dev)
        w = (torch.randn(a.Nout, a.K, generator=g) / a.K ** 0.5).bfloat16().to(dev)
        act = {"none": L.ACT_NONE, "gelu": L.ACT_GELU_TANH, "silu": L.ACT_SILU}[a.act]
        resid = torch.randn(a.M, a.Nout, generator=g).to(dev) if a.resid else None
        odt = torch.float32 if (a.out == "f32" or a.resid) else torch.bfloat16
        out = torch.empty(a.M, a.Nout, dtype=odt, device=dev)
        fn = lambda: ops.gemm(x, w, act=act, resid=resid, alpha=0.5, out_dtype=odt, impl=L.GEMM_TCGEN05, out=out)
        flops = 2.0 * a.M * a.Nout * a.K
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"{a.op} {vars(a)}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
