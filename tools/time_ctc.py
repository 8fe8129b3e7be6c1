#!/usr/bin/env python
"""Time the CTC loss kernels (forward alpha, backward beta+grad) at a BASELINE config's size, next to
torch's own CUDA ctc_loss (ATen — what the reference calls on a GPU)."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lcasr_b200
from lcasr_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=45000); ap.add_argument("--B", type=int, default=1)
ap.add_argument("--V", type=int, default=4096); ap.add_argument("--frac", type=float, default=0.3)
ap.add_argument("--bwd", action="store_true"); ap.add_argument("--quick", action="store_true", help="only the wavefront forward (kernel experiments)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
S = int(a.frac * a.N)
g = torch.Generator().manual_seed(0)
lp = torch.randn(a.B, a.N, a.V, generator=g).log_softmax(-1).to(dev)
tgt = torch.randint(0, a.V - 1, (a.B, S), generator=g).to(dev)
il = torch.full((a.B,), a.N, dtype=torch.int32, device=dev)
tl = torch.full((a.B,), S, dtype=torch.int64, device=dev)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
nll, _ = ops.ctc_loss_fwd(lp, tgt, il, tl, a.V - 1)
if a.quick:
    print(f"N={a.N} S={S} B={a.B} TB={os.environ.get('LCASR_CTC_WF_TB', '16')} DBG={os.environ.get('LCASR_CTC_WF_DBG', '0')}: "
          f"wavefront fwd {t(lambda: ops.ctc_loss_fwd(lp, tgt, il, tl, a.V - 1), 5):.2f} ms, nll {nll.tolist()[:2]}")
    sys.exit(0)
ref = torch.nn.functional.ctc_loss(lp.transpose(0, 1), tgt, il.long(), tl, blank=a.V - 1, reduction="none")
print(f"N={a.N} S={S} B={a.B}: ours nll {nll.tolist()} torch {ref.tolist()}")
from lcasr_b200 import _lib as L
nll_p = torch.empty_like(nll)
def plain():
    L.call("lcasr_ctc_loss_fwd", L.ptr(lp), a.B, a.N, a.V, L.ptr(tgt), S, L.ptr(il), L.ptr(tl), a.V - 1, L.ptr(nll_p), None, L.current_stream())
plain(); torch.cuda.synchronize()
print(f"  wavefront applies: {ops.ctc_wavefront_applies(a.B, a.N, S, False)}; bit-identical to the plain recursion: {torch.equal(nll, nll_p)}")
print(f"  plain (one CTA / cluster) fwd {t(plain):.2f} ms")
print(f"  ours fwd {t(lambda: ops.ctc_loss_fwd(lp, tgt, il, tl, a.V - 1)):.2f} ms ; torch(ATen CUDA) fwd {t(lambda: torch.nn.functional.ctc_loss(lp.transpose(0, 1), tgt, il.long(), tl, blank=a.V - 1, reduction='none')):.2f} ms")
if a.bwd:
    def ours_bwd():
        nl, al = ops.ctc_loss_fwd(lp, tgt, il, tl, a.V - 1, keep_alpha=True)
        return ops.ctc_loss_bwd(lp, tgt, il, tl, a.V - 1, nl, torch.ones_like(nl), al)
    def torch_bwd():
        x = lp.clone().requires_grad_(True)
        torch.nn.functional.ctc_loss(x.transpose(0, 1), tgt, il.long(), tl, blank=a.V - 1, reduction="sum").backward()
        return x.grad
    g1, g2 = ours_bwd(), torch_bwd()
    print(f"  grad max-abs diff vs torch {(g1 - g2).abs().max().item():.3e}")
    print(f"  ours fwd+bwd {t(ours_bwd, 2):.2f} ms ; torch fwd+bwd {t(torch_bwd, 2):.2f} ms")
