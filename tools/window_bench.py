"""Local vs full attention kernel time (tcgen05, bf16): python tools/window_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lcasr_b200 import ops

dev = torch.device("cuda", 0)
for (B, N, H, Dh, w) in [(1, 16384, 24, 32, 256), (1, 45056, 16, 128, 256), (1, 16384, 6, 128, 1024)]:
    q, k, v = (torch.randn(B, N, H, Dh, device=dev).bfloat16() for _ in range(3))
    res = {"B": B, "N": N, "H": H, "Dh": Dh, "window": w}
    for name, fn in (("full", lambda: ops.attention(q, k, v)), ("window", lambda: ops.attention_window(q, k, v, w, w))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_us"] = e0.elapsed_time(e1) / 10 * 1e3
    band = sum(min(N - 1, i + w) - max(0, i - w) + 1 for i in range(0, N, 64)) * 64  # ~ visible (query, key) pairs
    res["window_tflops_on_band"] = 4.0 * B * H * band * Dh / res["window_us"] / 1e6
    res["speedup"] = res["full_us"] / res["window_us"]
    print(json.dumps(res))
