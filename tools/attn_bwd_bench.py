#!/usr/bin/env python
"""Attention backward: flash-style kernels (lcasr_attention_bwd_flash) against the materialised form (P/dS kernel + three
batched GEMMs) — time, TFLOP/s (10*B*H*N^2*Dh useful FLOPs) and agreement.
  python tools/attn_bwd_bench.py [--shapes 8,2048,6,128 1,16384,6,128]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcasr_b200 import train_ops as T  # noqa: E402


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", nargs="*", default=["8,2048,6,128", "1,16384,6,128", "2,4096,12,64"])
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    for sh in a.shapes:
        B, N, H, Dh = (int(x) for x in sh.split(","))
        g = torch.Generator(device=dev).manual_seed(3)
        q, k, v = (torch.randn(B, N, H, Dh, generator=g, device=dev).bfloat16() for _ in range(3))
        do = torch.randn(B, N, H, Dh, generator=g, device=dev).bfloat16()
        out, lse2 = T.attention_train(q, k, v)
        o4 = out.view(B, N, H, Dh)
        flops = 10.0 * B * H * N * N * Dh
        res = {}
        for name, kw in (("flash", dict(flash=True)), ("materialised", dict(flash=False))):
            if name == "materialised" and H * N * N * 4 > (8 << 30):
                continue
            ms = timeit(lambda: T.attention_bwd(q, k, v, o4, do, lse2, **kw), a.iters)
            res[name] = T.attention_bwd(q, k, v, o4, do, lse2, **kw)
            print(f"B={B} N={N} H={H} Dh={Dh} {name:13s} {ms:8.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
        if len(res) == 2:
            for nm, x, y in zip(("dq", "dk", "dv"), res["flash"], res["materialised"]):
                print(f"   {nm}: max |flash - materialised| = {(x.float() - y.float()).abs().max().item():.3e} "
                      f"(max |.| {y.float().abs().max().item():.3e})")


if __name__ == "__main__":
    main()
