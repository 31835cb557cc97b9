"""TEST INFRASTRUCTURE -- numpy restatement of the conditional RQ-spline MAF the reference takes
from the third-party `zuko` package.  PARITY UNPINNED: zuko (pinned `zuko=1.2.0`, reference
env.yml:40; Dockerfile:7 installs probabilists/zuko@master; README.md:59-61 names the fork
valsdav/zuko) is neither vendored under /root/reference nor installed here, and the reference
holds no golden log_prob/sample values.  This file restates zuko's published algorithm
(zuko/nn.py MaskedMLP, zuko/flows/autoregressive.py MaskedAutoregressiveTransform / MAF,
zuko/flows/spline.py NSF / NCSF, zuko/transforms.py MonotonicRQSTransform /
AutoregressiveTransform / CircularShiftTransform, zuko/distributions.py DiagNormal / BoxUniform);
parity is anchored on
  * the reference's call sites (memflow/unfolding_flow/unfolding_flow.py:88-97,179,186;
    memflow/transfer_flow/transfer_flow_paper_AllPartons_btag.py:80-129,202-213;
    memflow/ttH/transfer_flow/custom_flows.py:10-127; memflow/transfer_flow/periodicNSF_gaussian.py),
  * the one structural known-answer the reference records: the module tree and parameter count
    (30 738) printed in notebooks/TestFlows_ImportanceSampling.ipynb cells 9 and 12,
  * self-consistency (inverse o forward = id, ladj = log|det autograd Jacobian|, normalisation).
Version-sensitive details (soft-clip `slope`, circular-shift inverse) are explicit parameters.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
Everything is float64 unless `dtype=np.float32` is passed (the reference's fp32 torch arithmetic).
"""
import math

import numpy as np

# univariate kinds
RQS = 0           # MonotonicRQSTransform(*phi, bound)
CIRCULAR_RQS = 1  # ComposedTransform(CircularShiftTransform(bound), MonotonicRQSTransform(*phi, bound))
PS_RQS = 2        # ComposedTransform(PSTransform(), MonotonicRQSTransform(*phi, bound=1))  (custom_flows.py:43-48)
# bases
DIAG_NORMAL = 0   # DiagNormal(loc, scale)
BOX_UNIFORM = 1   # BoxUniform(lower, upper)


# ------------------------------------------------------------------------------------------------
# masks: zuko.nn.MaskedMLP.__init__ and zuko.flows.autoregressive.MaskedAutoregressiveTransform.__init__
def autoregressive_adjacency(features, context, passes, order, total):
    """adjacency [features*total, features+context] and the effective (integer-divided) order."""
    if passes is None:
        passes = features
    order = np.arange(features) if order is None else np.asarray(order, dtype=np.int64)
    passes = min(max(int(passes), 1), features)
    order = order // int(math.ceil(features / passes))
    in_order = np.concatenate([order, np.full(context, -1, dtype=np.int64)])
    out_order = np.repeat(order, total)
    return out_order[:, None] > in_order[None, :], order, passes


def masked_mlp_masks(adjacency, hidden_features):
    """Boolean masks [out_l, in_l] of every MaskedLinear of zuko's MaskedMLP(adjacency, hidden)."""
    adjacency = np.asarray(adjacency, dtype=bool)
    uniq, inverse = np.unique(adjacency, axis=0, return_inverse=True)  # lexicographic, like torch.unique(dim=0)
    inverse = np.asarray(inverse).reshape(-1)
    a = uniq.astype(np.int64)
    precedence = (a @ a.T) == a.sum(-1)[None, :]
    masks = []
    indices = None
    n_layers = len(hidden_features) + 1
    for i, feats in enumerate(list(hidden_features) + [adjacency.shape[0]]):
        mask = precedence[:, indices] if i > 0 else uniq
        if (~mask).all():
            raise ValueError("The adjacency matrix leads to a null Jacobian.")
        if i < n_layers - 1:
            reachable = np.nonzero(mask.sum(-1))[0]
            indices = reachable[np.arange(feats) % len(reachable)]
            mask = mask[indices]
        else:
            mask = mask[inverse]
        masks.append(mask)
    return masks


# ------------------------------------------------------------------------------------------------
def masked_mlp(x, layers):
    """layers: list of (W [out,in], b [out], mask [out,in]); ReLU between layers (zuko default)."""
    h = x
    for li, (W, b, mask) in enumerate(layers):
        h = h @ (mask * W).T + b
        if li < len(layers) - 1:
            h = np.maximum(h, 0)
    return h


def _softclip(x, factor):
    return x / (1 + np.abs(factor * x))


def rqs_knots(w, h, d, bound, slope):
    """MonotonicRQSTransform.__init__: unconstrained (w [..,K], h [..,K], d [..,K-1]) -> knots."""
    dt = w.dtype
    if slope is not None:
        ls = dt.type(math.log(slope))
        w = w / (1 + np.abs(2 * w / ls))
        h = h / (1 + np.abs(2 * h / ls))
        d = d / (1 + np.abs(d / ls))

    def softmax(a):
        a = a - a.max(-1, keepdims=True)
        e = np.exp(a)
        return e / e.sum(-1, keepdims=True)

    pad0 = np.zeros(w.shape[:-1] + (1,), dtype=dt)
    W = np.concatenate([pad0, softmax(w)], -1)
    H = np.concatenate([pad0, softmax(h)], -1)
    D = np.concatenate([pad0, d, pad0], -1)
    bound = dt.type(bound)
    horizontal = bound * (2 * np.cumsum(W, -1, dtype=dt) - 1)
    vertical = bound * (2 * np.cumsum(H, -1, dtype=dt) - 1)
    return horizontal, vertical, np.exp(D)


def _bin(k, hor, ver, der):
    K = hor.shape[-1] - 1
    mask = (0 <= k) & (k < K)
    k = np.mod(k, K)
    take = lambda a, idx: np.take_along_axis(a, idx[..., None], -1)[..., 0]
    x0, x1 = take(hor, k), take(hor, k + 1)
    y0, y1 = take(ver, k), take(ver, k + 1)
    d0, d1 = take(der, k), take(der, k + 1)
    s = (y1 - y0) / (x1 - x0)
    return mask, k, x0, x1, y0, y1, d0, d1, s


def rqs_forward(x, hor, ver, der):
    """MonotonicRQSTransform.call_and_ladj -> (y, ladj, bin index, in-range mask)."""
    k = (hor < x[..., None]).sum(-1) - 1  # searchsorted
    mask, kk, x0, x1, y0, y1, d0, d1, s = _bin(k, hor, ver, der)
    z = mask * (x - x0) / (x1 - x0)
    y = y0 + (y1 - y0) * (s * z**2 + d0 * z * (1 - z)) / (s + (d0 + d1 - 2 * s) * z * (1 - z))
    jac = s**2 * (2 * s * z * (1 - z) + d0 * (1 - z) ** 2 + d1 * z**2) / (s + (d0 + d1 - 2 * s) * z * (1 - z)) ** 2
    return np.where(mask, y, x), mask * np.log(jac), kk, mask


def rqs_inverse(y, hor, ver, der):
    """MonotonicRQSTransform._inverse -> (x, bin index, mask)."""
    k = (ver < y[..., None]).sum(-1) - 1
    mask, kk, x0, x1, y0, y1, d0, d1, s = _bin(k, hor, ver, der)
    y_ = mask * (y - y0)
    a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - 2 * s)
    b = (y1 - y0) * d0 - y_ * (d0 + d1 - 2 * s)
    c = -s * y_
    z = 2 * c / (-b - np.sqrt(b**2 - 4 * a * c))
    x = x0 + z * (x1 - x0)
    return np.where(mask, x, y), kk, mask


def _wrap(x, bound):
    return np.mod(x + bound, 2 * bound) - bound  # CircularShiftTransform (torch `%` = floor mod)


class FlowSpec:
    """Plain container describing one flow in zuko's terms.

    transforms: list of dicts {"order": int[D] (already integer-divided), "passes": int,
                               "layers": [(W, b, mask), ...]}, in the order zuko applies them for
                               log_prob (data -> base).
    """

    def __init__(self, features, context, bins, transforms, bound=5.0, slope=1e-3, univariate=RQS,
                 base=DIAG_NORMAL, base_a=None, base_b=None):
        self.features, self.context, self.bins = features, context, bins
        self.transforms = transforms
        self.bound, self.slope, self.univariate = float(bound), slope, univariate
        self.base = base
        self.base_a = np.zeros(features) if base_a is None else np.asarray(base_a, dtype=np.float64)
        self.base_b = np.ones(features) if base_b is None else np.asarray(base_b, dtype=np.float64)


def _univariate_params(spec, t, x, c, dtype):
    inp = np.concatenate([x, c], -1) if spec.context > 0 else x
    layers = [(W.astype(dtype), b.astype(dtype), m.astype(dtype)) for (W, b, m) in t["layers"]]
    phi = masked_mlp(inp.astype(dtype), layers)
    K = spec.bins
    phi = phi.reshape(phi.shape[:-1] + (spec.features, 3 * K - 1))
    return rqs_knots(phi[..., :K], phi[..., K:2 * K], phi[..., 2 * K:], spec.bound, spec.slope)


def _uni_forward(spec, x, knots):
    ladj0 = 0.0
    if spec.univariate == CIRCULAR_RQS:
        x = _wrap(x, x.dtype.type(spec.bound))
    elif spec.univariate == PS_RQS:
        x = (x + 1) / 2
        ladj0 = math.log(0.5)
    y, ladj, k, mask = rqs_forward(x, *knots)
    return y, ladj + ladj0, k, mask


def _uni_inverse(spec, y, knots):
    x, k, mask = rqs_inverse(y, *knots)
    if spec.univariate == CIRCULAR_RQS:
        x = _wrap(x, x.dtype.type(spec.bound))
    elif spec.univariate == PS_RQS:
        x = 2 * x - 1
    return x


def base_log_prob(spec, z):
    if spec.base == DIAG_NORMAL:
        loc, scale = spec.base_a.astype(z.dtype), spec.base_b.astype(z.dtype)
        lp = -((z - loc) ** 2) / (2 * scale**2) - np.log(scale) - z.dtype.type(math.log(math.sqrt(2 * math.pi)))
        return lp.sum(-1)
    lo, hi = spec.base_a.astype(z.dtype), spec.base_b.astype(z.dtype)
    inside = (lo <= z) & (hi > z)  # torch.distributions.Uniform.log_prob
    with np.errstate(divide="ignore"):
        lp = np.log(inside.astype(z.dtype)) - np.log(hi - lo)
    return lp.sum(-1)


def flow_forward(spec, x, c, dtype=np.float64, return_debug=False):
    """x [N,D], c [N,C] -> (z, ladj [N]); optionally the bin indices / masks of every transform."""
    x = np.asarray(x, dtype=dtype)
    c = np.asarray(c, dtype=dtype) if spec.context > 0 else None
    ladj = np.zeros(x.shape[:-1], dtype=dtype)
    dbg = []
    for t in spec.transforms:
        knots = _univariate_params(spec, t, x, c, dtype)
        x, l, k, mask = _uni_forward(spec, x, knots)
        ladj = ladj + l.sum(-1)
        dbg.append((k, mask))
    return (x, ladj, dbg) if return_debug else (x, ladj)


def flow_log_prob(spec, x, c, dtype=np.float64):
    z, ladj = flow_forward(spec, x, c, dtype)
    return base_log_prob(spec, z) + ladj


def flow_inverse(spec, z, c, dtype=np.float64):
    """base -> data: transforms in reverse, each inverted with `passes` fixed-point sweeps
    (zuko AutoregressiveTransform._inverse)."""
    y = np.asarray(z, dtype=dtype)
    c = np.asarray(c, dtype=dtype) if spec.context > 0 else None
    for t in reversed(spec.transforms):
        x = np.zeros_like(y)
        for _ in range(t["passes"]):
            knots = _univariate_params(spec, t, x, c, dtype)
            x = _uni_inverse(spec, y, knots)
        y = x
    return y


def base_sample_from_noise(spec, noise):
    """Deterministic base sample from caller-supplied noise: standard normals for DiagNormal,
    uniforms in [0,1) for BoxUniform (identical-input-noise parity protocol)."""
    if spec.base == DIAG_NORMAL:
        return spec.base_a + spec.base_b * noise
    return spec.base_a + (spec.base_b - spec.base_a) * noise


# ------------------------------------------------------------------------------------------------
def random_spec(features, context, bins, n_transforms, hidden, seed, passes=None, randperm=False,
                bound=5.0, slope=1e-3, univariate=RQS, base=DIAG_NORMAL, base_a=None, base_b=None,
                weight_scale=1.0):
    """zuko-style random-init flow (nn.Linear default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)))."""
    rng = np.random.default_rng(seed)
    total = 3 * bins - 1
    transforms = []
    for i in range(n_transforms):
        if randperm:
            order = rng.permutation(features)
        else:
            order = np.arange(features) if i % 2 == 0 else np.arange(features)[::-1]
        adj, eff_order, eff_passes = autoregressive_adjacency(features, context, passes, order, total)
        masks = masked_mlp_masks(adj, hidden)
        layers = []
        for m in masks:
            fan_in = m.shape[1]
            k = weight_scale / math.sqrt(fan_in)
            layers.append((rng.uniform(-k, k, size=m.shape), rng.uniform(-k, k, size=m.shape[0]), m))
        transforms.append({"order": eff_order, "passes": eff_passes, "layers": layers})
    return FlowSpec(features, context, bins, transforms, bound, slope, univariate, base, base_a, base_b)
