"""depthwise Conv1d kernels of the conv module (training forms at cfg-5 size, inference form at cfg-3 / cfg-2 size), CUDA events.
Run twice to compare: LCASR_DWCONV_LEGACY=1 selects the register-window kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcasr_b200 import ops, train_ops as T
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2] * 1e3


tag = "legacy" if os.environ.get("LCASR_DWCONV_LEGACY") else "tile"
for (B, N, d) in [(8, 2048, 768), (1, 16384, 768), (16, 2048, 768), (1, 45000, 2048)]:
    x = torch.randn(B, N, d, generator=g).to(torch.bfloat16).to(dev)
    dy = torch.randn(B, N, d, generator=g).to(torch.bfloat16).to(dev)
    w, b = (torch.randn(d, 9, generator=g) * 0.3).to(dev), (torch.randn(d, generator=g) * 0.1).to(dev)
    mean, std = torch.zeros(d, device=dev), torch.ones(d, device=dev)
    dw, db = torch.zeros(d, 9, device=dev), torch.zeros(d, device=dev)
    mb = B * N * d * 2 / 1e6
    res = {
        "eval (dw+BRN+SiLU)": timeit(lambda: ops.dwconv_brn_silu(x, w, b, mean, std, std, mean)),
        "fwd+stats": timeit(lambda: T.dwconv1d_fwd(x, w, b, stats=True)),
        "bwd_data": timeit(lambda: T.dwconv1d_bwd_data(dy, w)),
        "bwd_weight": timeit(lambda: T.dwconv1d_bwd_weight_(x, dy, dw, db)),
    }
    print(f"[{tag}] B={B} N={N} d={d} ({mb:.0f} MB per tensor): " + "  ".join(f"{k} {v:.1f} us" for k, v in res.items()))
