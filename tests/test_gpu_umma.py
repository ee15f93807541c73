"""tcgen05 / TMEM building blocks: the self-test GEMM against a bf16 matmul."""
import pytest
import torch

import ertdiff_b200 as eb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 32), (32, 128), (128, 128), (64, 96)])
def test_umma_selftest_gemm(cuda_dev, N, K):
    g = torch.Generator().manual_seed(N * 1000 + K)
    a = torch.randn(128, K, generator=g).to(cuda_dev)
    b = torch.randn(N, K, generator=g).to(cuda_dev)
    got = eb.debug_umma_gemm(a, b)
    ref = a.bfloat16().double() @ b.bfloat16().double().t()          # exact products of bf16 operands
    err = (got.double() - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item(), err                  # fp32 accumulation order only
    # structure check: a one-hot A row picks a row of B exactly
    a1 = torch.zeros(128, K, device=cuda_dev)
    a1[torch.arange(128), torch.arange(128) % K] = 1.0
    got = eb.debug_umma_gemm(a1, b)
    assert torch.equal(got, b.bfloat16().float()[:, torch.arange(128) % K].t().contiguous())
