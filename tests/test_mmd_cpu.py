"""SURVEY 8(f)-4 (MMD loss): the numpy oracle against golden values of the UNMODIFIED reference
(memflow/unfolding_flow/mmd_loss.py, run by tools/make_golden_mmd.py), and the host-side argument handling of the
product mirror (which has no CPU path)."""
import os

import numpy as np
import pytest
import torch

from oracle import mmd_oracle


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "mmd.npz"))


def test_oracle_value_matches_reference(gold):
    for B, D in gold["cases"]:
        tag = f"B{B}_D{D}"
        x, y = gold[tag + "_x"], gold[tag + "_y"]
        for kernel in ("multiscale", "rbf"):
            ref = float(gold[f"{tag}_{kernel}_value"])
            got = float(mmd_oracle.mmd(x, y, kernel))
            # same operations in the same order; only the BLAS summation order inside the Gram products may differ
            assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (tag, kernel, got, ref)
            ref32 = float(gold[f"{tag}_{kernel}_value_f32"])
            got32 = float(mmd_oracle.mmd(x.astype(np.float32), y.astype(np.float32), kernel))
            assert abs(got32 - ref32) <= 2e-5 * max(1.0, abs(ref)), (tag, kernel, got32, ref32)


def test_oracle_gradient_matches_reference_autograd(gold):
    for B, D in gold["cases"]:
        tag = f"B{B}_D{D}"
        x, y = gold[tag + "_x"], gold[tag + "_y"]
        for kernel in ("multiscale", "rbf"):
            gx, gy = mmd_oracle.mmd_grad(x, y, kernel)
            scale = max(1e-30, np.abs(gold[f"{tag}_{kernel}_gx"]).max(), np.abs(gold[f"{tag}_{kernel}_gy"]).max())
            assert np.abs(gx - gold[f"{tag}_{kernel}_gx"]).max() <= 1e-10 * max(scale, 1.0), (tag, kernel)
            assert np.abs(gy - gold[f"{tag}_{kernel}_gy"]).max() <= 1e-10 * max(scale, 1.0), (tag, kernel)


def test_oracle_unknown_kernel_is_zero():
    x = np.random.default_rng(0).normal(size=(5, 3))
    assert mmd_oracle.mmd(x, x + 1.0, "laplace") == 0.0


def test_product_mirror_refuses_cpu_tensors_and_checks_shapes():
    from memflow_msci_project_b200 import _cabi
    from memflow_msci_project_b200.unfolding_flow.mmd_loss import MMD
    with pytest.raises(_cabi.MemflowB200Error):
        MMD(torch.rand(4, 3), torch.rand(4, 3), "multiscale", torch.device("cpu"), torch.float32)
