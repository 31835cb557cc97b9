"""GPU: tcgen05 building blocks (swizzled operands, descriptors, TMEM, MMA, commit, read-back) in
isolation, against torch matmul in fp64 on the same fp32 inputs."""
import pytest
import torch

from memflow_msci_project_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (64, 128), (80, 128), (128, 128), (128, 144), (128, 256), (256, 64)])
@pytest.mark.parametrize("mode,tol", [(0, 2e-6), (2, 2e-3), (1, 1.6e-2), (3, 2e-6), (4, 2e-6)])
def test_probe_gemm(cuda, K, N, mode, tol):
    g = torch.Generator().manual_seed(K * 1000 + N)
    a = torch.randn(128, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    out = _cabi.tc_probe_gemm(a, w, mode)
    ref = (a.double() @ w.double().t())
    scale = (a.double().abs() @ w.double().abs().t())  # sum of |terms|: the natural error scale of a dot product
    err = ((out.double() - ref).abs() / scale).max().item()
    print(f"K={K} N={N} mode={mode}: max err / sum|terms| = {err:.2e}")
    assert err < tol
