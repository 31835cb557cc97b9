"""SURVEY 8(f)-4 (periodic-phi regression loss): the numpy oracle against golden values of the UNMODIFIED reference
function (scripts/run_model_labframe_sampling.py:65-77, executed by tools/make_golden_losses.py with torch's own
HuberLoss / MSELoss), and the argument handling of the product mirror (which has no CPU path)."""
import os

import numpy as np
import pytest
import torch

from oracle import losses_oracle

CASES = [("huber1", "huber", 1.0, "mean"), ("huber03", "huber", 0.3, "mean"), ("mse", "mse", 0.0, "mean"),
         ("huber1_none", "huber", 1.0, "none"), ("huber1_sum", "huber", 1.0, "sum")]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


def upstream(n, reduction, dtype):
    """the weights tools/make_golden_losses.py multiplied the loss with before backward()"""
    if reduction == "none":
        return np.linspace(0.5, 1.5, n).astype(dtype) if dtype == np.float64 else \
            torch.linspace(0.5, 1.5, n, dtype=torch.float32).numpy()
    return (1.0 / n if reduction == "mean" else 1.0) * np.ones(n, dtype)


@pytest.mark.parametrize("tag,dtype,tol", [("f64", np.float64, 1e-13), ("f32", np.float32, 2e-6)])
def test_oracle_matches_reference(gold, tag, dtype, tol):
    inp, target = gold["inp"].astype(dtype), gold["target"].astype(dtype)
    for name, kind, delta, reduction in CASES:
        v, g = losses_oracle.loss_fn_periodic(inp, target, kind, delta, reduction)
        ref = gold[f"{name}_{tag}_value"]
        assert np.abs(np.asarray(v, np.float64) - ref).max() <= tol * max(1.0, np.abs(ref).max()), (name, tag)
        gt = g * upstream(inp.size, reduction, dtype)
        assert np.abs(gt - gold[f"{name}_{tag}_gtarget"]).max() <= tol * 10, (name, tag)
        assert np.abs(-gt - gold[f"{name}_{tag}_ginp"]).max() <= tol * 10, (name, tag)


def test_product_mirror_refuses_cpu_tensors_and_unknown_losses():
    from memflow_msci_project_b200 import _cabi
    from memflow_msci_project_b200.unfolding_flow.losses import loss_fn_periodic
    with pytest.raises(_cabi.MemflowB200Error):
        loss_fn_periodic(torch.rand(4), torch.rand(4), torch.nn.HuberLoss(), torch.device("cpu"))
