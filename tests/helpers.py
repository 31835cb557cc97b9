"""Shared comparison helpers of the parity tests."""
import numpy as np


def wrap_diff(a, b):
    """|a-b| for quantities living on the unit circle [0,1) (phi columns wrap 0 <-> 1)."""
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    return np.minimum(d, 1.0 - d)


def momenta_rel_err(p, p_ref):
    """max over components of |dP| / E_ref, per particle: the scale the 1e-5 gate refers to."""
    p = np.asarray(p, dtype=np.float64)
    p_ref = np.asarray(p_ref, dtype=np.float64)
    E = np.abs(p_ref[..., 0:1])
    return np.max(np.abs(p - p_ref) / np.maximum(E, 1e-300), axis=-1)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


GOLDEN_RAMBO = ["n3_Htt", "n4_HttG", "n5_HttGG", "n6_tt_decay", "n8_ttH_decay"]


def spec_from_module(flow):
    """Build the oracle's FlowSpec from a product flow module (parameters copied to numpy)."""
    from oracle import flow_oracle as fo
    core = flow.core_transforms()
    t0 = core[0]
    transforms = []
    for t in core:
        layers = [(lin.weight.detach().double().cpu().numpy(), lin.bias.detach().double().cpu().numpy(),
                   lin.mask.detach().cpu().numpy().astype(bool)) for lin in t.hyper.linears()]
        transforms.append({"order": t.order.cpu().numpy(), "passes": t.passes, "layers": layers})
    a, b = flow.base_tensors()
    return fo.FlowSpec(t0.features, t0.context, t0.bins, transforms, bound=t0.bound, slope=t0.slope,
                       univariate=t0.univariate, base=flow.base_kind(), base_a=a.double().cpu().numpy(),
                       base_b=b.double().cpu().numpy())
