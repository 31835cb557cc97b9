"""CPU: the flow oracle against the reference's only recorded known-answer and its own invariants;
the product's host logic (mask construction, module tree, state-dict layout) against the oracle."""
import os

import numpy as np
import pytest
import torch

from memflow_msci_project_b200 import flows
from memflow_msci_project_b200.flows import masks as pmasks
from oracle import flow_oracle as fo


def test_notebook_known_answer_parameter_count():
    """notebooks/TestFlows_ImportanceSampling.ipynb cells 7, 9, 12: NSF(7, transforms=4, bins=8,
    hidden=[32]*3, BoxUniform base_args, passes=7) prints MaskedLinear 7->32->32->32->161 and has
    30738 trainable parameters (4 x 7681 hyper-network + 14 base)."""
    spec = fo.random_spec(7, 0, 8, 4, [32] * 3, seed=0, passes=7, randperm=True, bound=1.0)
    shapes = [m.shape for (_, _, m) in spec.transforms[0]["layers"]]
    assert shapes == [(32, 7), (32, 32), (32, 32), (161, 32)]
    assert sum(W.size + b.size for t in spec.transforms for (W, b, m) in t["layers"]) + 14 == 30738
    flow = flows.NSF(7, transforms=4, bins=8, hidden_features=[32] * 3, randperm=True, base=flows.BoxUniform,
                     base_args=[-torch.ones(7), torch.ones(7)], univariate_kwargs={"bound": 1}, passes=7)
    assert sum(p.numel() for p in flow.parameters() if p.requires_grad) == 30738
    text = repr(flow)
    assert "MaskedLinear(in_features=7, out_features=32" in text and "out_features=161" in text
    keys = list(flow.state_dict().keys())
    assert keys[:4] == ["transforms.0.order", "transforms.0.hyper.0.weight", "transforms.0.hyper.0.bias",
                        "transforms.0.hyper.0.mask"]
    assert "base._0" in keys and "base._1" in keys


@pytest.mark.parametrize("D,C,K,H,passes", [(3, 65, 16, [128] * 4, 3), (10, 16, 40, [512] * 2, 2), (1, 66, 16, [256] * 3, 1),
                                            (7, 0, 8, [32] * 3, 7), (2, 65, 16, [128] * 4, 2), (12, 32, 35, [128] * 4, 2)])
def test_product_masks_equal_oracle_masks(D, C, K, H, passes):
    g = torch.Generator().manual_seed(D * 100 + C)
    order = torch.randperm(D, generator=g)
    adj_t, o_t, p_t = pmasks.autoregressive_adjacency(D, C, passes, order, 3 * K - 1)
    adj_n, o_n, p_n = fo.autoregressive_adjacency(D, C, passes, order.numpy(), 3 * K - 1)
    assert np.array_equal(adj_t.numpy(), adj_n) and np.array_equal(o_t.numpy(), o_n) and p_t == p_n
    for a, b in zip(pmasks.mlp_masks(adj_t, H), fo.masked_mlp_masks(adj_n, H)):
        assert np.array_equal(a.numpy(), b)


def test_masks_are_autoregressive():
    """The composed mask connectivity must respect the order: output block of feature i may depend on
    feature j only if order[i] > order[j] (and on every context input)."""
    D, C, K, H = 5, 3, 4, [16, 16]
    order = np.array([3, 0, 4, 1, 2])
    adj, eff, _ = fo.autoregressive_adjacency(D, C, None, order, 3 * K - 1)
    masks = fo.masked_mlp_masks(adj, H)
    conn = masks[0].astype(np.int64)
    for m in masks[1:]:
        conn = (m.astype(np.int64) @ conn > 0).astype(np.int64)
    for i in range(D):
        rows = conn[i * (3 * K - 1):(i + 1) * (3 * K - 1)]
        for j in range(D):
            assert rows[:, j].all() == (eff[i] > eff[j]) and rows[:, j].any() == (eff[i] > eff[j])
        assert rows[:, D:].all()


@pytest.mark.parametrize("univariate", [fo.RQS, fo.CIRCULAR_RQS, fo.PS_RQS])
def test_oracle_inverse_and_jacobian(univariate):
    """inverse(forward(x)) = x and ladj = log|det J| by finite differences of the oracle's forward."""
    bound = 1.0 if univariate == fo.PS_RQS else 2.0
    spec = fo.random_spec(3, 4, 6, 3, [32, 32], seed=3, passes=3, bound=bound, univariate=univariate,
                          weight_scale=3.0)
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.9, 0.9, size=(64, 3)) * (1.0 if univariate == fo.PS_RQS else bound)
    c = rng.normal(size=(64, 4))
    z, ladj = fo.flow_forward(spec, x, c)
    assert np.abs(fo.flow_inverse(spec, z, c) - x).max() < 1e-9
    eps = 1e-6
    J = np.zeros((64, 3, 3))
    for j in range(3):
        dx = np.zeros(3); dx[j] = eps
        J[:, :, j] = (fo.flow_forward(spec, x + dx, c)[0] - fo.flow_forward(spec, x - dx, c)[0]) / (2 * eps)
    logdet = np.log(np.abs(np.linalg.det(J)))
    assert np.abs(logdet - ladj).max() < 1e-5


def test_oracle_density_normalised_1d():
    spec = fo.random_spec(1, 2, 8, 3, [16, 16], seed=5, bound=1.0, base=fo.BOX_UNIFORM, base_a=[-1.0], base_b=[1.0],
                          weight_scale=2.0)
    xs = np.linspace(-1, 1, 400001)[:, None]
    c = np.tile(np.array([[0.3, -1.2]]), (xs.shape[0], 1))
    p = np.exp(fo.flow_log_prob(spec, xs, c))
    assert p.max() > 5 and abs(np.trapezoid(p, xs[:, 0]) - 1) < 1e-6


def test_box_uniform_support_and_identity_outside_bound():
    spec = fo.random_spec(2, 0, 4, 2, [8], seed=1, bound=1.0, base=fo.BOX_UNIFORM, base_a=[-1, -1], base_b=[1, 1])
    x = np.array([[0.2, 0.1], [1.5, 0.0], [0.0, -3.0]])
    z, ladj = fo.flow_forward(spec, x, None)
    assert z[1, 0] == 1.5 and z[2, 1] == -3.0  # identity outside [-bound, bound]
    lp = fo.flow_log_prob(spec, x, None)
    assert np.isfinite(lp[0]) and lp[1] == -np.inf and lp[2] == -np.inf


@pytest.mark.parametrize("ctor,kw", [
    (flows.NSF, dict(features=3, context=5, transforms=3, bins=6, hidden_features=[32, 32], univariate_kwargs={"bound": 1.5})),
    (flows.NCSF_gaussian, dict(features=2, context=3, bins=5, transforms=2, hidden_features=[16])),
    (flows.PSNSF, dict(features=2, context=0, bins=7, transforms=2, hidden_features=[24])),
])
def test_differentiable_restatement_matches_oracle(ctor, kw):
    """The torch restatement that provides the (interim) backward pass computes the same function as the oracle."""
    from helpers import spec_from_module
    from memflow_msci_project_b200.flows import _autograd as ag
    torch.manual_seed(3)
    flow = ctor(**kw).double()
    with torch.no_grad():
        for t in flow.core_transforms():
            t.hyper.linears()[-1].weight.mul_(3.0)
    spec = spec_from_module(flow)
    g = torch.Generator().manual_seed(5)
    D, C = flow.features, flow.context
    x = (torch.rand(200, D, generator=g, dtype=torch.float64) * 2 - 1) * 0.9
    c = torch.randn(200, C, generator=g, dtype=torch.float64) if C else None
    z, ladj = ag.eager_forward(flow, x, c)
    z_ref, ladj_ref = fo.flow_forward(spec, x.numpy(), c.numpy() if c is not None else None)
    assert np.abs(z.detach().numpy() - z_ref).max() < 1e-12 and np.abs(ladj.detach().numpy() - ladj_ref).max() < 1e-11
    xb = ag.eager_inverse(flow, torch.from_numpy(z_ref), c)
    assert np.abs(xb.detach().numpy() - fo.flow_inverse(spec, z_ref, c.numpy() if c is not None else None)).max() < 1e-10


def test_gemm_auto_resolution_table():
    """Which kernel a call takes under gemm="auto" (flows/modules.py::MAF.effective_gemm) -- host logic only."""
    import torch
    from memflow_msci_project_b200 import flows
    from memflow_msci_project_b200.flows import modules as fm
    mk = lambda **kw: flows.NSF(**kw)
    tfa = mk(features=3, context=65, transforms=2, bins=16, hidden_features=[128] * 4)
    uf = mk(features=10, context=16, transforms=2, bins=40, hidden_features=[512] * 2)
    tf1d = mk(features=1, context=66, transforms=2, bins=16, hidden_features=[256] * 3)
    huge = mk(features=2, context=4, transforms=1, bins=8, hidden_features=[1024])
    manybins = mk(features=2, context=4, transforms=1, bins=50, hidden_features=[64])
    f32, f64 = torch.float32, torch.float64
    assert tfa.gemm == fm.GEMM_AUTO                                         # the default
    assert tfa.effective_gemm(f32) == fm.GEMM_TC_3XF16
    assert tfa.effective_gemm(f32, sampling=True) == fm.GEMM_TC_3XF16
    assert tfa.effective_gemm(f32, ctx_group=8) == fm.GEMM_TC_3XF16         # context grouping: in the tcgen05 kernel too (r2)
    noclip = mk(features=3, context=65, transforms=2, bins=16, hidden_features=[128] * 4, univariate_kwargs={"slope": None})
    assert noclip.effective_gemm(f32) == fm.GEMM_MMA_3XF16                  # no soft clip: not in the tcgen05 kernel
    assert tfa.effective_gemm(f64) == fm.GEMM_EXACT
    assert uf.effective_gemm(f32) == fm.GEMM_MMA_3XF16 and uf.effective_gemm(f32, sampling=True) == fm.GEMM_MMA_3XF16
    assert tf1d.effective_gemm(f32) == fm.GEMM_MMA_3XF16
    assert huge.effective_gemm(f32) == fm.GEMM_EXACT                        # wider than 512
    assert manybins.effective_gemm(f32) == fm.GEMM_EXACT                    # 3K-1 > 128
    noclip.gemm = fm.GEMM_TC_3XF16                                          # asked for explicitly: falls to the parity anchor
    assert noclip.effective_gemm(f32) == fm.GEMM_EXACT


def test_gemm_constants_match_header():
    import re
    from memflow_msci_project_b200.flows import modules as fm
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "memflow_b200.h")).read()
    vals = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define (MF_GEMM_\w+) (\d+)", hdr)}
    assert vals == {"MF_GEMM_EXACT": fm.GEMM_EXACT, "MF_GEMM_TC_3XF16": fm.GEMM_TC_3XF16,
                    "MF_GEMM_TC_BF16": fm.GEMM_TC_BF16, "MF_GEMM_MMA_3XF16": fm.GEMM_MMA_3XF16}


@pytest.mark.parametrize("D,passes,randperm", [(3, None, False), (10, 2, True), (7, 3, True), (5, 5, False)])
def test_scheduled_inverse_equals_literal_sweeps_in_the_oracle(D, passes, randperm):
    """The argument behind mf_flow_set_sample_schedule, checked on the CPU oracle alone: updating in sweep s only the
    features of order group s gives exactly the samples of zuko's literal `passes` full sweeps."""
    spec = fo.random_spec(D, 4, 8, 3, [32, 32], seed=5, passes=passes, randperm=randperm, weight_scale=3.0)
    rng = np.random.default_rng(1)
    z = rng.standard_normal((200, D))
    c = rng.standard_normal((200, 4))
    literal = fo.flow_inverse(spec, z, c)
    y = z.copy()
    for t in reversed(spec.transforms):
        x = np.zeros_like(y)
        order = np.asarray(t["order"])
        for s in range(t["passes"]):
            knots = fo._univariate_params(spec, t, x, c, np.float64)
            xn = fo._uni_inverse(spec, y, knots)
            x = np.where(order[None, :] == s, xn, x)      # only the group that becomes final in this sweep
        y = x
    assert np.array_equal(literal, y)
