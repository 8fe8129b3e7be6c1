"""CPU: the numerical argument behind conv0 on the tensor cores (csrc/subsample_tc.cu).  The fp32 spectrogram patch and the
fp32 weights are split into bf16 hi + lo parts; the K = 32 product
    [x_hi | x_lo | x_hi | 1 | 1 | 0 0 0] . [w_hi | w_hi | w_lo | b_hi | b_lo | 0 0 0]      (w, b pre-scaled by 1/2)
with fp32 accumulation is (conv0 + bias) / 2 up to the dropped x_lo*w_lo term (~2^-16 relative), and
SiLU(x) = h + h*tanh(h) with h = x / 2.  bf16 products of bf16 values are exact in fp32, so the statement below (bf16-rounded
operands, fp32 accumulate) is what the tensor core computes up to the summation order."""
import torch


def _split(t):
    hi = t.to(torch.bfloat16).to(torch.float32)
    lo = (t - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def test_hi_lo_split_product_is_fp32_grade():
    torch.manual_seed(0)
    P, C = 4096, 64
    patch = torch.randn(P, 9) * 3.0                       # standardised log-mel values
    w = torch.randn(C, 9) / 3.0
    b = 0.1 * torch.randn(C)
    ref = (patch.double() @ w.double().T + b.double()) * 0.5
    xh, xl = _split(patch)
    wh, wl = _split(0.5 * w)                              # exact: the split of w/2 is half the split of w
    bh, bl = _split(0.5 * b)
    ones = torch.ones(P, 1)
    A = torch.cat([xh, xl, xh, ones, ones, torch.zeros(P, 3)], 1)
    B = torch.cat([wh, wh, wl, bh[:, None], bl[:, None], torch.zeros(C, 3)], 1)
    assert A.shape[1] == 32 and B.shape[1] == 32
    acc = A @ B.T                                         # fp32 accumulate of exact bf16 x bf16 products
    err = (acc.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err < 2.0 ** -14 * scale, (err, scale)         # far below the bf16 rounding (2^-9) of the stored activation
    # a single bf16 pass (what a plain bf16 GEMM would give) is two orders worse: the split is what keeps fp32 grade
    single = (xh @ wh.T + bh)
    assert (single.double() - ref).abs().max().item() > 20 * err
    # positions outside the conv0 range get an all-zero A row: accumulator exactly 0 and SiLU(0) = 0 (no select needed)
    assert float((torch.zeros(1, 32) @ B.T).abs().max()) == 0.0


def test_silu_from_half_argument():
    x = torch.linspace(-12, 12, 4001, dtype=torch.float64)
    h = 0.5 * x
    assert torch.allclose(h + h * torch.tanh(h), torch.nn.functional.silu(x), atol=1e-12)
