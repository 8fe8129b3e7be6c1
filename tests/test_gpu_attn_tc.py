"""GPU: the tcgen05 flash-attention kernel against fp32 SDPA on the same bf16 operands
(attention.py:541 is the reference's CPU path; its CUDA path is flash-attn 2 in fp16)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import report

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


CASES = [(32, 128, 1, 1), (32, 33, 2, 2), (32, 300, 3, 1), (64, 257, 2, 1), (128, 128, 1, 1), (128, 125, 1, 1),
         (128, 640, 2, 2), (32, 2048, 4, 1), (128, 2048, 2, 1), (32, 4100, 2, 1)]


@pytest.mark.parametrize("vt", [True, False])
@pytest.mark.parametrize("Dh,N,H,B", CASES)
def test_attention_tcgen05(cuda_device, Dh, N, H, B, vt):
    from lcasr_b200 import ops, _lib as L
    # scaled so that some rows have sharp maxima (exercises the lazy O rescale) and some are flat
    q = (_rand(B, N, H, Dh, seed=1) * torch.linspace(0.2, 3.0, N)[None, :, None, None]).bfloat16()
    k = _rand(B, N, H, Dh, seed=2).bfloat16()
    v = _rand(B, N, H, Dh, seed=3).bfloat16()
    qc, kc, vc = q.to(cuda_device), k.to(cuda_device), v.to(cuda_device)
    ref = F.scaled_dot_product_attention(qc.float().transpose(1, 2), kc.float().transpose(1, 2), vc.float().transpose(1, 2))
    ref = ref.transpose(1, 2).reshape(B, N, H * Dh)
    if vt:
        Npad = (N + 127) // 128 * 128
        vin = torch.zeros(B, H, Dh, Npad, dtype=torch.bfloat16, device=cuda_device)
        vin[..., :N] = vc.permute(0, 2, 3, 1)
    else:
        vin = vc
    got = ops.attention(qc, kc, vin, v_transposed=vt, impl=L.ATTN_TCGEN05)
    torch.cuda.synchronize()
    err = (got.float() - ref).abs().max().item()
    report(test="attn_tc", Dh=Dh, N=N, H=H, B=B, vt=vt, max_abs=err)
    assert torch.isfinite(got.float()).all()
    assert err < 2e-2, f"tcgen05 attention mismatch {err}"
    simt = ops.attention(qc, kc, vc, impl=L.ATTN_SIMT)
    assert (got.float() - simt.float()).abs().max().item() < 2e-2
