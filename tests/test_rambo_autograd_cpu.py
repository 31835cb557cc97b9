"""RAMBO autograd (generateKinematics_batch(..., requires_grad=True)): the differentiable restatement that runs the
backward is checked on CPU tensors against gradients produced by the reference's own autograd graph
(tests/golden/rambo_grad_*.npz, written by tools/make_golden_rambo.py) and against finite differences."""
import os

import numpy as np
import pytest
import torch

from memflow_msci_project_b200.phasespace import _autograd as ag
from memflow_msci_project_b200.phasespace.phasespace import PhaseSpace

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _gen(masses):
    n = len(masses)
    ps = PhaseSpace(13000, [21, 21], [0] * n, final_masses=torch.tensor(masses), dev=torch.device("cpu"))
    gen = ps.generator
    return gen, gen._desc("fwd", gen.masses_t, (-1, -1, -1))


@pytest.mark.parametrize("case", ["n3_Htt", "n4_HttG"])
def test_reference_graph_gradient(case):
    g = np.load(os.path.join(GOLDEN, f"rambo_grad_{case}.npz"))
    gen, d = _gen([float(m) for m in g["masses"]])
    r = torch.from_numpy(g["r"]).requires_grad_(True)
    mom, w, x1, x2 = ag.eager_forward(gen, d, r, full_derivative=False)
    loss = ((mom * torch.from_numpy(g["cot_momenta"])).sum() + (w * torch.from_numpy(g["cot_weight"])).sum()
            + (x1 * torch.from_numpy(g["cot_x1"])).sum() + (x2 * torch.from_numpy(g["cot_x2"])).sum())
    loss.backward()
    ref = torch.from_numpy(g["grad_r"])
    n = gen.n_final
    assert torch.all(ref[:, : n - 2] == 0) and torch.all(r.grad[:, : n - 2] == 0)   # quirk: no gradient through the root
    scale = ref.abs().max(0).values.clamp_min(1e-30)
    assert ((r.grad - ref).abs() / scale).max().item() < 1e-6                      # fp32 volume rounding: ~1e-7


def test_full_derivative_matches_finite_differences():
    gen, d = _gen([125.25, 172.5, 172.5, 1e-5])
    g = torch.Generator().manual_seed(5)
    r0 = torch.rand(16, 10, generator=g, dtype=torch.float64).clamp_(0.05, 0.95)
    cm = torch.randn(16, 6, 4, generator=g, dtype=torch.float64)

    def f(r):
        mom, w, x1, x2 = ag.eager_forward(gen, d, r, full_derivative=True)
        return (mom * cm).sum(dim=(1, 2)) + torch.log(w) + x1 - 2 * x2

    r = r0.clone().requires_grad_(True)
    f(r).sum().backward()
    eps = 1e-6
    for col in range(10):
        dr = torch.zeros_like(r0)
        dr[:, col] = eps
        fd = (f(r0 + dr) - f(r0 - dr)) / (2 * eps)
        rel = (fd - r.grad[:, col]).abs() / (fd.abs() + 1e-3)
        assert rel.max().item() < 1e-5, (col, rel.max().item())
