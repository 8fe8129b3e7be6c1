"""GPU: the tcgen05 / TMEM / TMA GEMM against fp32 matmul of the same bf16 operands."""
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import report

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


SHAPES = [(128, 128, 64), (128, 256, 64), (256, 256, 128), (66, 192, 64), (333, 128, 320), (1000, 4096, 256),
          (2048, 768, 3072), (4096, 2304, 768), (130, 32, 32), (20000, 256, 256), (513, 72, 40), (16384, 768, 768)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tcgen05_plain(cuda_device, M, N, K):
    from lcasr_b200 import ops, _lib as L
    a = _rand(M, K, seed=1).bfloat16()
    w = (_rand(N, K, seed=2) / math.sqrt(K)).bfloat16()
    ref = a.float().to(cuda_device) @ w.float().to(cuda_device).T
    got = ops.gemm(a.to(cuda_device), w.to(cuda_device), out_dtype=torch.float32, impl=L.GEMM_TCGEN05)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    report(test="gemm_tc", M=M, N=N, K=K, max_abs=err)
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), f"tcgen05 GEMM mismatch {err}"
    gb = ops.gemm(a.to(cuda_device), w.to(cuda_device), out_dtype=torch.bfloat16, impl=L.GEMM_TCGEN05)
    assert (gb.float() - ref).abs().max().item() < 2 ** -7 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(333, 256, 128), (2048, 768, 3072), (1000, 4096, 256)])
def test_gemm_tcgen05_epilogues(cuda_device, M, N, K):
    from lcasr_b200 import ops, _lib as L
    a = _rand(M, K, seed=3).bfloat16().to(cuda_device)
    w = (_rand(N, K, seed=4) / math.sqrt(K)).bfloat16().to(cuda_device)
    bias = _rand(N, seed=5).to(cuda_device)
    resid = _rand(M, N, seed=6).to(cuda_device)
    ref = a.float() @ w.float().T + bias
    got = ops.gemm(a, w, bias=bias, act=L.ACT_GELU_TANH, out_dtype=torch.bfloat16, impl=L.GEMM_TCGEN05)
    assert (got.float() - F.gelu(ref, approximate="tanh")).abs().max().item() < 3e-2
    got = ops.gemm(a, w, bias=bias, act=L.ACT_SILU, out_dtype=torch.float32, impl=L.GEMM_TCGEN05)
    assert (got - F.silu(ref)).abs().max().item() < 2e-3
    r = resid.clone()
    got = ops.gemm(a, w, bias=bias, resid=r, alpha=0.5, impl=L.GEMM_TCGEN05, out=r)
    assert (got - (resid + 0.5 * ref)).abs().max().item() < 2e-3
    # agreement with the independent SIMT kernel on identical operands
    simt = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, impl=L.GEMM_SIMT)
    tc = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, impl=L.GEMM_TCGEN05)
    assert (simt - tc).abs().max().item() < 1e-3
