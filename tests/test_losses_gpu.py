"""SURVEY 8(f)-4 on the B200: ``loss_fn_periodic`` (mf_periodic_loss, csrc/periodic.cu) against the goldens of the
unmodified reference function (tests/golden/losses.npz) and the numpy oracle.

Tolerances: per-element loss and derivative are a handful of exactly-rounded operations in the input dtype in the same
order as torch's, so they are compared at a few ulp of the dtype (fp64 4e-15 relative to max(1, |v|), fp32 1e-6); the mean /
sum is accumulated in fp64 here and in the tensor dtype by torch: 1e-13 (fp64), 2e-6 (fp32) relative.
"""
import os

import numpy as np
import pytest
import torch

from oracle import losses_oracle
from test_losses_cpu import CASES, upstream

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


def _loss_fn(kind, delta, reduction):
    return torch.nn.MSELoss(reduction=reduction) if kind == "mse" else torch.nn.HuberLoss(delta=delta, reduction=reduction)


@pytest.mark.parametrize("tag,dtype,elem_tol,sum_tol", [("f64", torch.float64, 4e-15, 1e-13), ("f32", torch.float32, 1e-6, 2e-6)])
def test_matches_reference_golden(cuda, gold, tag, dtype, elem_tol, sum_tol):
    from memflow_msci_project_b200.unfolding_flow.losses import loss_fn_periodic
    npdt = np.float64 if dtype == torch.float64 else np.float32
    for name, kind, delta, reduction in CASES:
        inp = torch.from_numpy(gold["inp"]).to(dtype).to(cuda).requires_grad_(True)
        target = torch.from_numpy(gold["target"]).to(dtype).to(cuda).requires_grad_(True)
        before = inp.detach().clone()
        v = loss_fn_periodic(inp, target, _loss_fn(kind, delta, reduction), cuda)
        assert v.dtype == dtype
        ref = gold[f"{name}_{tag}_value"]
        assert v.shape == ref.shape
        tol = elem_tol if reduction == "none" else sum_tol
        assert np.abs(v.detach().cpu().numpy().astype(np.float64) - ref).max() <= tol * max(1.0, np.abs(ref).max()), (name, tag)
        w = torch.from_numpy(upstream(inp.numel(), reduction, npdt)).to(cuda) if reduction == "none" else None
        (v * w).sum().backward() if w is not None else v.backward()
        assert np.abs(inp.grad.cpu().numpy() - gold[f"{name}_{tag}_ginp"]).max() <= elem_tol * 4, (name, tag)
        assert np.abs(target.grad.cpu().numpy() - gold[f"{name}_{tag}_gtarget"]).max() <= elem_tol * 4, (name, tag)
        assert torch.equal(inp.detach(), before)


def test_large_batch_against_oracle_2d_and_edge_cases(cuda):
    from memflow_msci_project_b200.unfolding_flow.losses import loss_fn_periodic
    rng = np.random.default_rng(11)
    a, t = 3.0 * rng.normal(size=(1000003,)), rng.uniform(-np.pi, np.pi, size=(1000003,))
    want, _ = losses_oracle.loss_fn_periodic(a, t, "huber", 0.7, "mean")
    x, y = torch.from_numpy(a).to(cuda), torch.from_numpy(t).to(cuda)
    fn = torch.nn.HuberLoss(delta=0.7)
    got = loss_fn_periodic(x, y, fn, cuda)
    assert abs(got.item() - want) <= 1e-13 * max(1.0, abs(want))
    assert loss_fn_periodic(x, y, fn, cuda).item() == got.item()               # fixed-order reduction
    el, _ = losses_oracle.loss_fn_periodic(a[:600].reshape(200, 3), t[:600].reshape(200, 3), "mse", 0.0, "none")
    got2 = loss_fn_periodic(x[:600].reshape(200, 3), y[:600].reshape(200, 3), torch.nn.MSELoss(reduction="none"), cuda)
    assert got2.shape == (200, 3) and np.abs(got2.cpu().numpy() - el).max() <= 1e-14
    assert torch.isnan(loss_fn_periodic(x[:0], y[:0], fn, cuda))
    with pytest.raises(TypeError):
        loss_fn_periodic(x, y, torch.nn.L1Loss(), cuda)
    with pytest.raises(RuntimeError):
        loss_fn_periodic(x, y[:5], fn, cuda)


def test_nan_inputs_propagate_like_torch(cuda):
    """the reference has no guard: a NaN prediction gives a NaN loss element and a NaN gradient there, nothing else"""
    from memflow_msci_project_b200.unfolding_flow.losses import loss_fn_periodic
    inp = torch.tensor([0.3, float("nan"), -4.0, 2.0], dtype=torch.float64, device=cuda, requires_grad=True)
    target = torch.tensor([0.1, 0.2, 3.0, float("nan")], dtype=torch.float64, device=cuda)
    v = loss_fn_periodic(inp, target, torch.nn.HuberLoss(delta=0.5, reduction="none"), cuda)
    v.sum().backward()
    nan = torch.tensor([False, True, False, True], device=cuda)
    assert torch.equal(torch.isnan(v), nan) and torch.equal(torch.isnan(inp.grad), nan)
    want, g = losses_oracle.loss_fn_periodic(np.array([0.3, -4.0]), np.array([0.1, 3.0]), "huber", 0.5, "none")
    assert np.abs(v.detach().cpu().numpy()[[0, 2]] - want).max() <= 1e-15
    assert np.abs(inp.grad.cpu().numpy()[[0, 2]] + g).max() <= 1e-15
    assert torch.isnan(loss_fn_periodic(inp.detach(), target, torch.nn.HuberLoss(delta=0.5), cuda))
