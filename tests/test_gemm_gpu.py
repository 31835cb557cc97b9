"""mf_gemm_nt_3xf16 (tcgen05, fp16 hi/lo split operands) against fp64 matmul: the GEMM shapes of the hyper-network
backward -- forward recompute with bias + ReLU, input gradient, split-K weight gradient with operand scaling."""
import pytest
import torch

from memflow_msci_project_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _err(c, ref, a, b):
    """error relative to sum |a||b| (the natural scale of a dot product)"""
    scale = (a.double().abs() @ b.double().abs().t()).clamp_min(1e-300)
    return ((c.double() - ref).abs() / scale).max().item()


@pytest.mark.parametrize("m,n,k,bias,relu", [(1000, 128, 68, True, True), (4099, 141, 128, True, False), (300, 128, 141, False, False),
                                              (128, 16, 64, False, False), (77, 256, 200, True, True), (513, 300, 96, True, False),
                                              # K = 64 / 128 with several row tiles: the persistent kernel (resident weights,
                                              # bulk-copied A K-blocks), incl. a ragged last tile and more tiles than SMs
                                              (5000, 128, 128, True, True), (2001, 68, 128, False, False), (1000, 128, 64, True, False),
                                              (148 * 128 * 2 + 300, 141, 128, True, True)])
def test_gemm_nt_matches_fp64(cuda, m, n, k, bias, relu):
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).to(cuda)
    b = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda)
    bv = torch.randn(n, generator=g).to(cuda) if bias else None
    c = torch.full((m, n), float("nan"), device=cuda)
    _cabi.gemm_nt(a, b, c, bias=bv, relu=relu)
    ref = a.double() @ b.double().t()
    if bias:
        ref = ref + bv.double()
    if relu:
        ref = ref.clamp_min(0)
    assert torch.isfinite(c).all()
    assert _err(c, ref, a, b) < 2e-6


def test_gemm_nt_split_k_with_operand_scaling(cuda):
    """weight-gradient shape: a short, very wide product whose first operand is far below the fp16 range"""
    g = torch.Generator().manual_seed(5)
    rows, out, inn = 70001, 141, 128
    dz = (torch.randn(rows, out, generator=g) * 1e-7).to(cuda)
    h = torch.randn(rows, inn, generator=g).to(cuda)
    a, b = dz.t().contiguous(), h.t().contiguous()
    ref = a.double() @ b.double().t()
    amax = a.abs().amax()
    scale = torch.exp2(torch.floor(8.0 - torch.log2(amax))).reshape(1)
    c = torch.zeros(out, inn, device=cuda)
    _cabi.gemm_nt(a, b, c, k_split=96, scale_a=scale)
    assert _err(c, ref, a, b) < 2e-6
    # without the scaling the fp16 split flushes the operand: the scaled result must be much closer
    c0 = torch.zeros(out, inn, device=cuda)
    _cabi.gemm_nt(a, b, c0, k_split=96)
    assert _err(c0, ref, a, b) > 10 * _err(c, ref, a, b)


def test_gemm_nt_strided_views(cuda):
    """leading dimensions: column slices of wider matrices as operands and result"""
    g = torch.Generator().manual_seed(9)
    big_a = torch.randn(515, 200, generator=g).to(cuda)
    big_b = torch.randn(96, 160, generator=g).to(cuda)
    big_c = torch.zeros(515, 128, device=cuda)
    a, b, c = big_a[:, 8:136], big_b[:, 32:160], big_c[:, 16:112]
    _cabi.gemm_nt(a, b, c)
    ref = a.double() @ b.double().t()
    assert _err(c, ref, a, b) < 2e-6
    assert (big_c[:, :16] == 0).all() and (big_c[:, 112:] == 0).all()


@pytest.mark.parametrize("rows,out,inn", [(70001, 141, 128), (5000, 128, 68), (333, 16, 255)])
def test_gemm_tn_weight_gradient_with_relu_mask_and_bias_gradient(cuda, rows, out, inn):
    """dW = (dY * (Y > 0))^T X and db = column sums of the masked dY in one launch, straight from row-major operands."""
    g = torch.Generator().manual_seed(rows)
    dy = (torch.randn(rows, out, generator=g) * 1e-7 * torch.exp(2 * torch.randn(rows, 1, generator=g))).to(cuda)
    y = torch.randn(rows, out, generator=g).clamp_min(0).to(cuda)
    x = torch.randn(rows, inn, generator=g).to(cuda)
    dz = dy.double() * (y > 0)
    ref_w, ref_b = dz.t() @ x.double(), dz.sum(0)
    scale = torch.exp2(torch.floor(14.0 - torch.log2(dy.abs().amax()))).reshape(1)
    # (the tensor core accumulates in fp32 with truncation: a long K range in ONE accumulator picks up a bias that grows
    # with the number of MMAs -- ~2e-4 for 70 000 rows --, which is why the weight gradient is cut into short K ranges
    # whose partial sums are added with round-to-nearest atomics)
    for k_split in (37, 296):
        gw = torch.zeros(out, inn, device=cuda)
        gb = torch.zeros(out, device=cuda)
        _cabi.gemm_tn(dy, x, gw, k_split=k_split, mask=y, colsum=gb, scale_a=scale)
        assert ((gw.double() - ref_w).norm() / ref_w.norm()).item() < 3e-5
        assert ((gb.double() - ref_b).norm() / ref_b.norm()).item() < 3e-5


@pytest.mark.parametrize("rows,out,inn", [(3001, 141, 128), (3001, 128, 128), (40000, 128, 68)])
def test_gemm_nt_with_relu_mask(cuda, rows, out, inn):
    g = torch.Generator().manual_seed(3)
    dy = torch.randn(rows, out, generator=g).to(cuda)
    y = torch.randn(rows, out, generator=g).clamp_min(0).to(cuda)
    wt = (torch.randn(inn, out, generator=g) / out ** 0.5).to(cuda)      # W^T [in, out]
    gx = torch.empty(rows, inn, device=cuda)
    _cabi.gemm_nt(dy, wt, gx, mask=y)
    ref = (dy.double() * (y > 0)) @ wt.double().t()
    assert ((gx.double() - ref).norm() / ref.norm()).item() < 2e-6


@pytest.mark.parametrize("rows,out,inn", [(70001, 128, 128), (40000, 128, 68), (5000, 64, 128), (333, 128, 16),
                                          (64 * 700, 128, 141)])   # 141 columns: rows that are not 16-byte multiples, whole K-blocks only
def test_gemm_tn_bulk_copied_operands(cuda, rows, out, inn):
    """contiguous operands without a mask: the kernel that takes its K-blocks by bulk copy (tail K-block, several splits)"""
    g = torch.Generator().manual_seed(rows + out)
    dz = (torch.randn(rows, out, generator=g) * 1e-7 * torch.exp(2 * torch.randn(rows, 1, generator=g))).to(cuda)
    x = torch.randn(rows, inn, generator=g).to(cuda)
    ref_w, ref_b = dz.double().t() @ x.double(), dz.double().sum(0)
    scale = torch.exp2(torch.floor(14.0 - torch.log2(dz.abs().amax()))).reshape(1)
    for k_split in (7, 296):
        gw = torch.zeros(out, inn, device=cuda)
        gb = torch.zeros(out, device=cuda)
        _cabi.gemm_tn(dz, x, gw, k_split=k_split, colsum=gb, scale_a=scale)
        assert ((gw.double() - ref_w).norm() / ref_w.norm()).item() < 3e-5
        assert ((gb.double() - ref_b).norm() / ref_b.norm()).item() < 3e-5


@pytest.mark.parametrize("rows,out,inn", [(3001, 128, 128), (3001, 141, 128), (700, 128, 68)])
def test_gemm_nt_output_mask(cuda, rows, out, inn):
    """input gradient leaving already multiplied by the ReLU derivative of the layer below (its input x > 0)"""
    g = torch.Generator().manual_seed(11)
    dy = torch.randn(rows, out, generator=g).to(cuda)
    wt = (torch.randn(inn, out, generator=g) / out ** 0.5).to(cuda)
    x = torch.randn(rows, inn, generator=g).clamp_min(0).to(cuda)
    gx = torch.empty(rows, inn, device=cuda)
    _cabi.gemm_nt(dy, wt, gx, out_mask=x)
    ref = (dy.double() @ wt.double().t()) * (x > 0)
    assert ((gx.double() - ref).norm() / ref.norm()).item() < 2e-6
    assert (gx[x <= 0] == 0).all()
