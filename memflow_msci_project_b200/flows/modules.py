"""``nn.Module`` surface of the conditional spline flows, shaped like the zuko objects MEMFlow uses.

What the reference does with these objects (b):
  ``flow = zuko.flows.NSF(features=D, context=C, transforms=T, bins=K, hidden_features=[H]*L,
  randperm=..., passes=..., base=..., base_args=[a, b], univariate_kwargs={"bound": B})``
  (memflow/unfolding_flow/unfolding_flow.py:88-97, transfer_flow_paper_AllPartons_btag.py:80-129),
  then ``flow(c).log_prob(x)``, ``flow(c).rsample((S,))`` / ``.sample((S,))``, ``flow(c).transform(x)``,
  ``flow.transforms`` (list-like), ``flow.base()``, ``.float()/.double()/.to()``, state dicts.

Parameters keep zuko's layout (``transforms.{t}.hyper.{2l}.{weight,bias,mask}``, ``transforms.{t}.order``,
``base._0/_1``) so reference checkpoints load; the compute is ONE fused sm_100a kernel per call through
the C ABI (include/memflow_b200.h: mf_flow_log_prob / mf_flow_sample) on a packed copy of the weights
that is rebuilt whenever a parameter changes.  There is no torch/CPU fallback.
"""
import ctypes
import math

import torch
import torch.nn as nn

from .. import _cabi
from . import masks as _masks

MF_MAX_HIDDEN_LAYERS = 8
MF_MAX_FEATURES = 64

UNIV_RQS, UNIV_CIRCULAR_RQS, UNIV_PS_RQS = 0, 1, 2
BASE_DIAG_NORMAL, BASE_BOX_UNIFORM = 0, 1
GEMM_EXACT, GEMM_TC_3XF16, GEMM_TC_BF16, GEMM_MMA_3XF16 = 0, 1, 2, 3
# tensor paths pack the hidden units sorted by autoregressive degree (block-triangular masked weights: the kernels skip
# the all-zero blocks); the function computed is unchanged
SORTED_GEMMS = (GEMM_TC_3XF16, GEMM_TC_BF16, GEMM_MMA_3XF16)
GEMM_AUTO = -1  # resolved per call: tcgen05 3xFP16 when the shape fits, mma.sync 3xFP16 up to width 512, else exact


class FlowDesc(ctypes.Structure):
    """ctypes mirror of ``struct mf_flow_desc``."""

    _fields_ = [
        ("features", ctypes.c_int32), ("context", ctypes.c_int32), ("bins", ctypes.c_int32),
        ("transforms", ctypes.c_int32), ("n_hidden", ctypes.c_int32),
        ("hidden", ctypes.c_int32 * MF_MAX_HIDDEN_LAYERS),
        ("passes", ctypes.c_int32), ("univariate", ctypes.c_int32), ("base", ctypes.c_int32),
        ("ctx_group", ctypes.c_int32),
        ("bound", ctypes.c_double), ("slope", ctypes.c_double),
        ("base_a", ctypes.c_double * MF_MAX_FEATURES), ("base_b", ctypes.c_double * MF_MAX_FEATURES),
    ]


# ------------------------------------------------------------------------------------------------
# distributions (zuko.distributions.DiagNormal / BoxUniform are thin torch.distributions wrappers)
class DiagNormal(torch.distributions.Independent):
    def __init__(self, loc=0.0, scale=1.0, ndims=1):
        super().__init__(torch.distributions.Normal(torch.as_tensor(loc), torch.as_tensor(scale)), ndims)

    KIND = BASE_DIAG_NORMAL

    def expand(self, batch_shape, new=None):
        new = self._get_checked_instance(DiagNormal, new)
        return super().expand(batch_shape, new)


class BoxUniform(torch.distributions.Independent):
    def __init__(self, lower=0.0, upper=1.0, ndims=1):
        super().__init__(torch.distributions.Uniform(torch.as_tensor(lower), torch.as_tensor(upper),
                                                     validate_args=False), ndims)

    KIND = BASE_BOX_UNIFORM

    def expand(self, batch_shape, new=None):
        new = self._get_checked_instance(BoxUniform, new)
        return super().expand(batch_shape, new)


class Unconditional(nn.Module):
    """zuko.flows.core.Unconditional / UnconditionalDistribution: ``meta(*args)`` with the positional
    tensors registered as buffers (``buffer=True``) or parameters, under the names ``_0``, ``_1``..."""

    def __init__(self, meta, *args, buffer=False, **kwargs):
        super().__init__()
        self.meta = meta
        for i, arg in enumerate(args):
            arg = torch.as_tensor(arg)
            if buffer:
                self.register_buffer(f"_{i}", arg)
            else:
                self.register_parameter(f"_{i}", nn.Parameter(arg))
        self.kwargs = kwargs

    def tensors(self):
        return [t for n, t in sorted(list(self.named_buffers()) + list(self.named_parameters()))]

    def forward(self, c=None):
        return self.meta(*self.tensors(), **self.kwargs)

    def extra_repr(self):
        return getattr(self.meta, "__name__", repr(self.meta))


UnconditionalDistribution = Unconditional


# ------------------------------------------------------------------------------------------------
class MaskedLinear(nn.Linear):
    """zuko.nn.MaskedLinear: ``F.linear(x, mask * weight, bias)`` with a fixed boolean ``mask`` buffer.
    Only holds the parameters here; the product is formed inside the fused kernel."""

    def __init__(self, adjacency):
        super().__init__(adjacency.shape[1], adjacency.shape[0])
        self.register_buffer("mask", adjacency.to(torch.bool))

    def forward(self, x):  # pragma: no cover - the hot path never calls this
        raise _cabi.MemflowB200Error("MaskedLinear is evaluated inside the fused flow kernel only")


class MaskedMLP(nn.Sequential):
    """zuko.nn.MaskedMLP layout: MaskedLinear at even indices, ReLU in between."""

    def __init__(self, adjacency, hidden_features=(64, 64)):
        layers = []
        for mask in _masks.mlp_masks(adjacency, hidden_features):
            layers.extend([MaskedLinear(mask), nn.ReLU()])
        super().__init__(*layers[:-1])

    def linears(self):
        return [m for m in self if isinstance(m, MaskedLinear)]


class MaskedAutoregressiveTransform(nn.Module):
    """zuko.flows.autoregressive.MaskedAutoregressiveTransform with a rational-quadratic-spline
    univariate: hyper-network ``(x | c) -> D x (3K-1)`` spline parameters."""

    def __init__(self, features, context=0, passes=None, order=None, bins=8, hidden_features=(64, 64),
                 univariate=UNIV_RQS, bound=5.0, slope=1e-3):
        super().__init__()
        self.features, self.context, self.bins = features, context, bins
        self.univariate, self.bound, self.slope = univariate, float(bound), slope
        self.total = 3 * bins - 1
        adjacency, eff_order, self.passes = _masks.autoregressive_adjacency(features, context, passes,
                                                                            order, self.total)
        self.register_buffer("order", eff_order)
        self.hyper = MaskedMLP(adjacency, hidden_features)
        self.hidden_features = tuple(int(h) for h in hidden_features)

    def extra_repr(self):
        kinds = {UNIV_RQS: "MonotonicRQSTransform", UNIV_CIRCULAR_RQS: "CircularShift+MonotonicRQSTransform",
                 UNIV_PS_RQS: "PSTransform+MonotonicRQSTransform"}
        return f"(base): {kinds[self.univariate]}(bins={self.bins})\n(order): {self.order.tolist()}"


class SimpleAffineTransform(nn.Module):
    """Element-wise affine map of the interval [a_in, b_in] onto [a_out, b_out] (the zuko fork's
    ``SimpleAffineTransform``; notebooks/TestFlows_ImportanceSampling.ipynb cell 7 puts one in front
    of the spline transforms).  Evaluated with torch element-wise ops around the fused kernel."""

    def __init__(self, a_in, b_in, a_out, b_out):
        super().__init__()
        for n, v in (("a_in", a_in), ("b_in", b_in), ("a_out", a_out), ("b_out", b_out)):
            self.register_buffer(n, torch.as_tensor(v, dtype=torch.get_default_dtype()))

    def scale(self):
        return (self.b_out - self.a_out) / (self.b_in - self.a_in)

    def call_and_ladj(self, x):
        s = self.scale()
        return self.a_out + (x - self.a_in) * s, torch.log(torch.abs(s)).sum(-1).expand(x.shape[:-1])

    def inv(self, y):
        return self.a_in + (y - self.a_out) / self.scale()


# ------------------------------------------------------------------------------------------------
class SigmoidAffine:
    """``torch.sigmoid(x * scale + shift)`` applied to the samples INSIDE the sampling kernel (a19):
    ``flow(c).rsample((1,), post=SigmoidAffine(scale_ps, mean_ps))`` replaces the reference's
    ``torch.sigmoid(flow(c).rsample((1,)) * scale_ps + mean_ps)`` (scripts/run_model_labframe_sampling.py:426-427)."""

    def __init__(self, scale=1.0, shift=0.0):
        self.scale, self.shift = scale, shift

    def torch(self, x):
        return torch.sigmoid(x * torch.as_tensor(self.scale).to(x) + torch.as_tensor(self.shift).to(x))

    def xmap(self, features):
        return _cabi.make_xmap(_cabi.MF_XMAP_SIGMOID_AFFINE, self.scale, self.shift, features)


class LogitAffine:
    """``(torch.logit(x, eps) - mean) / std`` applied to the data INSIDE the log_prob kernel (a19): the bridge of
    memflow/unfolding_flow/unfolding_flow.py:170-171.  Like there it is a data transformation (no ladj term)."""

    def __init__(self, mean=0.0, std=1.0, eps=5e-5):
        self.mean, self.std, self.eps = mean, std, eps

    def torch(self, x):
        return (torch.logit(x, eps=self.eps) - torch.as_tensor(self.mean).to(x)) / torch.as_tensor(self.std).to(x)

    def xmap(self, features):
        return _cabi.make_xmap(_cabi.MF_XMAP_LOGIT_AFFINE, self.std, self.mean, features, eps=self.eps)


class _Transform:
    """What ``flow(c).transform`` returns: callable forward map with ``.inv`` and ``.call_and_ladj``."""

    def __init__(self, bound_flow):
        self._bf = bound_flow

    def __call__(self, x):
        return self._bf._forward(x, want_logprob=False)[0]

    def call_and_ladj(self, x):
        z, ladj, _ = self._bf._forward(x, want_logprob=False)
        return z, ladj

    def inv(self, z):
        return self._bf._inverse(z)


class NormalizingFlow:
    """``flow(c)``: zuko.distributions.NormalizingFlow bound to a context (a14-a16)."""

    def __init__(self, module, c, ctx_group=1):
        self._m = module
        self._c = c
        self._ctx_group = int(ctx_group)

    # -- properties used by the reference
    @property
    def transform(self):
        return _Transform(self)

    @property
    def base(self):
        d = self._m.base()
        return d if self._c is None else d.expand(self._c.shape[:-1])

    # -- helpers
    def _with_group(self, desc):
        if self._ctx_group <= 1:
            return desc
        d = FlowDesc.from_buffer_copy(desc)
        d.ctx_group = self._ctx_group
        return d

    def _split(self):
        pre, core, post, seen = [], [], [], False
        for t in self._m.transforms:
            if isinstance(t, MaskedAutoregressiveTransform):
                if post:
                    raise _cabi.MemflowB200Error("element-wise transforms between spline transforms are not supported")
                core.append(t)
                seen = True
            elif isinstance(t, SimpleAffineTransform):
                (post if seen else pre).append(t)
            else:
                raise _cabi.MemflowB200Error(f"unsupported transform module {type(t).__name__}")
        return pre, core, post

    def _rows(self, x):
        """Broadcast x [..., D] against the context [..., C] and flatten to rows."""
        D = self._m.features
        if x.shape[-1] != D:
            raise ValueError(f"expected last dimension {D}, got {tuple(x.shape)}")
        c = self._c
        if c is None:
            return x.reshape(-1, D), None, x.shape[:-1]
        if self._ctx_group > 1:
            xb = x.reshape(-1, D)
            cb = c.reshape(-1, c.shape[-1])
            if cb.shape[0] * self._ctx_group != xb.shape[0]:
                raise ValueError(f"ctx_group={self._ctx_group}: {cb.shape[0]} context rows do not cover {xb.shape[0]} x rows")
            return xb, cb, x.shape[:-1]
        batch = torch.broadcast_shapes(x.shape[:-1], c.shape[:-1])
        xb = x.expand(*batch, D).reshape(-1, D)
        cb = c.expand(*batch, c.shape[-1]).reshape(-1, c.shape[-1])
        return xb, cb, batch

    def _forward(self, x, want_logprob, debug=False, xmap=None):
        m = self._m
        pre, core, post = self._split()
        _cabi.require_cuda(x, "x")
        if torch.is_grad_enabled() and (x.requires_grad or (self._c is not None and self._c.requires_grad)
                                        or any(p.requires_grad for p in m.parameters())):
            from ._autograd import flow_log_prob_autograd
            if xmap is not None:   # under autograd the bridge is differentiable torch in front of the fused kernel
                x = xmap.torch(x)
            return flow_log_prob_autograd(self, x, want_logprob)
        if xmap is not None and pre:
            x, xmap = xmap.torch(x), None   # element-wise prefix transforms come after the bridge: apply it eagerly
        dtype = m.compute_dtype()
        ladj_pre = None
        x = x.to(dtype)
        for t in pre:
            x, l = t.call_and_ladj(x)
            ladj_pre = l if ladj_pre is None else ladj_pre + l
        xb, cb, batch = self._rows(x)
        xb = xb.contiguous()
        cb = cb.to(dtype).contiguous() if cb is not None else None
        N, D = xb.shape
        gemm = m.effective_gemm(dtype, ctx_group=self._ctx_group)
        desc, blob = m.packed(xb.device, dtype, gemm=gemm)
        desc = self._with_group(desc)
        z = torch.empty_like(xb)
        ladj = torch.empty(N, dtype=dtype, device=xb.device)
        need_lp = want_logprob and not post
        logprob = torch.empty(N, dtype=dtype, device=xb.device) if need_lp else None
        bins = inside = None
        if debug:
            bins = torch.empty((len(core), N, D), dtype=torch.int8, device=xb.device)
            inside = torch.empty((len(core), N, D), dtype=torch.uint8, device=xb.device)
        _cabi.flow_log_prob(desc, blob, xb, cb, logprob, z, ladj, bins, inside, gemm,
                            xmap=xmap.xmap(D) if xmap is not None else None)
        z = z.reshape(*batch, D)
        ladj = ladj.reshape(batch)
        if ladj_pre is not None:
            ladj = ladj + ladj_pre
            if logprob is not None:
                logprob = logprob + ladj_pre.reshape(-1)
        for t in post:
            z, l = t.call_and_ladj(z)
            ladj = ladj + l
        if want_logprob and logprob is None:
            logprob = self.base.log_prob(z).reshape(-1) + ladj.reshape(-1)
        if logprob is not None:
            logprob = logprob.reshape(batch)
        if debug:
            return z, ladj, logprob, bins.reshape(len(core), *batch, D), inside.reshape(len(core), *batch, D).bool()
        return z, ladj, logprob

    def _inverse(self, z, noise_is_base_sample=True, xmap=None):
        """x = T^-1(z) for z already distributed like the base."""
        m = self._m
        pre, core, post = self._split()
        _cabi.require_cuda(z, "z")
        dtype = m.compute_dtype()
        z = z.to(dtype)
        for t in reversed(post):
            z = t.inv(z)
        zb, cb, batch = self._rows(z)
        zb = zb.contiguous()
        cb = cb.to(dtype).contiguous() if cb is not None else None
        gemm = m.effective_gemm(dtype, sampling=True, ctx_group=self._ctx_group)
        desc, blob = m.packed(zb.device, dtype, identity_base=True, gemm=gemm)
        desc = self._with_group(desc)
        x = torch.empty_like(zb)
        fuse = xmap is not None and not pre
        _cabi.flow_sample(desc, blob, zb, cb, x, gemm, xmap=xmap.xmap(m.features) if fuse else None)
        x = x.reshape(*batch, m.features)
        for t in reversed(pre):
            x = t.inv(x)
        if xmap is not None and not fuse:
            x = xmap.torch(x)
        return x

    # -- distribution API
    def log_prob(self, x, pre=None):
        """``pre``: optional LogitAffine bridge evaluated inside the kernel on the data (extension, a19)."""
        return self._forward(x, want_logprob=True, xmap=pre)[2]

    def rsample(self, shape=(), post=None):
        """``post``: optional SigmoidAffine bridge evaluated inside the kernel on the samples (extension, a19)."""
        base = self.base
        z = base.rsample(torch.Size(shape)) if base.has_rsample else base.sample(torch.Size(shape))
        if torch.is_grad_enabled() and (z.requires_grad or (self._c is not None and self._c.requires_grad)
                                        or any(p.requires_grad for p in self._m.parameters())):
            from ._autograd import flow_rsample_autograd
            x = flow_rsample_autograd(self, z)
            return post.torch(x) if post is not None else x
        return self._inverse(z, xmap=post)

    def sample(self, shape=(), post=None):
        with torch.no_grad():
            base = self.base
            z = base.sample(torch.Size(shape))
            return self._inverse(z, xmap=post)

    def sample_from_noise(self, noise, post=None):
        """Deterministic sampling from caller-supplied noise [..., D] (standard normals for a
        DiagNormal base, uniforms in [0,1) for BoxUniform): the identical-noise parity protocol."""
        m = self._m
        a, b = m.base_tensors()
        a, b = a.to(noise), b.to(noise)
        z = a + b * noise if m.base_kind() == BASE_DIAG_NORMAL else a + (b - a) * noise
        with torch.no_grad():
            return self._inverse(z, xmap=post)

    def log_prob_debug(self, x):
        """(log_prob, z, ladj, bin indices [T,...,D] int8, in-range masks [T,...,D] bool)."""
        z, ladj, lp, bins, inside = self._forward(x, want_logprob=True, debug=True)
        return lp, z, ladj, bins, inside


# ------------------------------------------------------------------------------------------------
def _base_kind(meta):
    kind = getattr(meta, "KIND", None)
    if kind is None:
        name = getattr(meta, "__name__", "")
        kind = {"DiagNormal": BASE_DIAG_NORMAL, "BoxUniform": BASE_BOX_UNIFORM}.get(name)
    if kind is None:
        raise _cabi.MemflowB200Error(f"unsupported base distribution {meta!r} (DiagNormal / BoxUniform only)")
    return kind


class MAF(nn.Module):
    """zuko.flows.autoregressive.MAF with spline univariates = the NSF family.

    Accepts both constructor generations found in the reference: upstream
    zuko 1.x (``bins, transforms, randperm, passes, hidden_features``) and the valsdav fork
    (``base=, base_args=, univariate_kwargs={"bound": B}, dtype=``).
    """

    def __init__(self, features, context=0, transforms=3, randperm=False, bins=8, hidden_features=(64, 64),
                 passes=None, univariate=UNIV_RQS, bound=5.0, slope=1e-3, base=None, base_args=None,
                 univariate_kwargs=None, dtype=None, gemm="auto", **unused):
        super().__init__()
        if unused:
            raise TypeError(f"unsupported keyword arguments: {sorted(unused)}")
        if univariate_kwargs:
            bound = univariate_kwargs.get("bound", bound)
            slope = univariate_kwargs.get("slope", slope)
        self.features, self.context = int(features), int(context)
        orders = [torch.arange(features), torch.flipud(torch.arange(features))]
        self.transforms = nn.ModuleList([
            MaskedAutoregressiveTransform(features=features, context=context, passes=passes,
                                          order=torch.randperm(features) if randperm else orders[i % 2],
                                          bins=bins, hidden_features=hidden_features, univariate=univariate,
                                          bound=bound, slope=slope)
            for i in range(transforms)])
        if base is None:
            self.base = Unconditional(DiagNormal, torch.zeros(features), torch.ones(features), buffer=True)
        else:
            args = base_args if base_args is not None else [torch.zeros(features), torch.ones(features)]
            self.base = Unconditional(base, *args)  # the fork registers base_args as parameters
        self.gemm = {"exact": GEMM_EXACT, "tc_3xf16": GEMM_TC_3XF16, "tc_bf16": GEMM_TC_BF16, "mma_3xf16": GEMM_MMA_3XF16,
                     "auto": GEMM_AUTO}[gemm]
        self._blob_cache = {}
        self._desc_cache = {}
        self._register_load_state_dict_pre_hook(self._rename_legacy_keys)
        if dtype is not None:
            self.to(dtype)

    # zuko >= 1.0 nests the transforms: transform.transforms.{t}.*  ->  transforms.{t}.*
    @staticmethod
    def _rename_legacy_keys(state_dict, prefix, *args):
        for k in list(state_dict.keys()):
            if k.startswith(prefix + "transform.transforms."):
                state_dict[prefix + k[len(prefix) + len("transform."):]] = state_dict.pop(k)

    def forward(self, c=None, ctx_group=1):
        """``flow(c)``.  ``ctx_group=G`` (extension): ``c`` has one row per GROUP of G consecutive x rows -- the
        per-event context of a MEM integration sweep is read in place instead of being repeated G times."""
        if c is not None and c.shape[-1] != self.context:
            raise ValueError(f"expected context with last dimension {self.context}, got {tuple(c.shape)}")
        return NormalizingFlow(self, c, ctx_group)

    # -- descriptors / packed weights ---------------------------------------------------------
    def core_transforms(self):
        return [t for t in self.transforms if isinstance(t, MaskedAutoregressiveTransform)]

    def compute_dtype(self):
        p = next(self.core_transforms()[0].hyper.parameters())
        if p.dtype not in (torch.float32, torch.float64):
            raise _cabi.MemflowB200Error(f"flow parameters are {p.dtype}; float32/float64 only")
        return p.dtype

    def base_kind(self):
        return _base_kind(self.base.meta)

    def base_tensors(self):
        t = self.base.tensors()
        if len(t) != 2:
            raise _cabi.MemflowB200Error("base distribution must have exactly two tensor arguments")
        return t[0].detach(), t[1].detach()

    def make_desc(self, identity_base=False):
        core = self.core_transforms()
        t0 = core[0]
        for t in core[1:]:
            same = (t.features, t.context, t.bins, t.hidden_features, t.passes, t.univariate, t.bound, t.slope) == \
                   (t0.features, t0.context, t0.bins, t0.hidden_features, t0.passes, t0.univariate, t0.bound, t0.slope)
            if not same:
                raise _cabi.MemflowB200Error("all spline transforms of a flow must share one shape")
        if len(t0.hidden_features) > MF_MAX_HIDDEN_LAYERS or t0.features > MF_MAX_FEATURES:
            raise _cabi.MemflowB200Error("flow exceeds MF_MAX_HIDDEN_LAYERS / MF_MAX_FEATURES")
        d = FlowDesc()
        d.features, d.context, d.bins = t0.features, t0.context, t0.bins
        d.transforms, d.n_hidden = len(core), len(t0.hidden_features)
        for i, h in enumerate(t0.hidden_features):
            d.hidden[i] = h
        d.passes, d.univariate = t0.passes, t0.univariate
        d.bound = t0.bound
        d.slope = float(t0.slope) if t0.slope else 0.0
        if identity_base:  # sampling entry is fed base samples directly: z = 0 + 1 * noise
            d.base = BASE_DIAG_NORMAL
            for f in range(t0.features):
                d.base_a[f], d.base_b[f] = 0.0, 1.0
        else:
            d.base = self.base_kind()
            a, b = self.base_tensors()
            a, b = a.double().cpu(), b.double().cpu()
            for f in range(t0.features):
                d.base_a[f], d.base_b[f] = float(a[f]), float(b[f])
        return d

    def effective_gemm(self, dtype, sampling=False, ctx_group=1):
        """Conditioner arithmetic of one call.  fp64 flows run the CUDA-core kernel.  "auto": tcgen05 3xFP16 when the
        shape fits that kernel, else the warp-level MMA kernel up to hidden width 512, else the CUDA-core kernel.
        What the tcgen05 kernel does not take (splines without soft clip) goes to the MMA kernel under "auto" and to
        the CUDA-core kernel when a tcgen05 mode was asked for explicitly."""
        if dtype != torch.float32:
            return GEMM_EXACT
        t0 = self.core_transforms()[0]
        widths = [lin.weight.shape[0] for lin in t0.hyper.linears()[:-1]]
        mma_ok = max(widths) <= 512 and 3 * t0.bins - 1 <= 128
        gemm = self.gemm
        if gemm == GEMM_AUTO:
            kp = -(-t0.bins // 8) * 8
            tc_ok = max(widths) <= 128 and -(-self.context // 8) * 8 + self.features <= 128 and 3 * kp <= 128
            gemm = GEMM_TC_3XF16 if tc_ok else (GEMM_MMA_3XF16 if mma_ok else GEMM_EXACT)
        if gemm in (GEMM_TC_3XF16, GEMM_TC_BF16) and not t0.slope:
            return GEMM_MMA_3XF16 if (self.gemm == GEMM_AUTO and mma_ok) else GEMM_EXACT
        return gemm

    def _sched_words(self, device):
        """uint8 view of the [T][64] uint64 schedule words of mf_flow_set_sample_schedule (cached per device)."""
        core = self.core_transforms()
        # keyed on the identity / version of the order buffers: reading them back (a device -> host sync per
        # transform) happens only when they changed (construction, load_state_dict), not at every re-pack
        key = (str(device),) + tuple((t.order.data_ptr(), t.order._version) for t in core)
        cached = getattr(self, "_sched_cache", None)
        if cached is None or cached[0] != key:
            import numpy as np
            orders = tuple(tuple(int(v) for v in t.order.tolist()) for t in core)
            words = np.zeros((len(core), 64), dtype=np.uint64)
            for ti, od in enumerate(orders):
                for s in range(64):
                    m = 0
                    for f, g in enumerate(od):
                        if g == s:
                            m |= 1 << f
                    words[ti, s] = m if m else 1
            t = torch.from_numpy(words.view(np.uint8).reshape(-1).copy()).to(device)
            cached = (key, t)
            self._sched_cache = cached
        return cached[1]

    def packed(self, device, dtype, identity_base=False, gemm=None):
        """(desc, blob): the kernel-side copy of the weights, rebuilt when any parameter changed."""
        gemm = self.gemm if gemm is None else gemm
        core = self.core_transforms()
        params = []
        for t in core:
            for lin in t.hyper.linears():
                params += [lin.weight, lin.bias, lin.mask]
        key = (str(device), dtype, tuple((p.data_ptr(), p._version) for p in params))
        cached = self._blob_cache.get(gemm)
        if cached is None or cached[0] != key:
            for p in params:
                if p.device != device:
                    raise _cabi.MemflowB200Error(f"flow parameters live on {p.device}, data on {device}")
            desc = self.make_desc(identity_base=True)
            ws = [p.detach().to(dtype).contiguous() for p in params[0::3]]
            bs = [p.detach().to(dtype).contiguous() for p in params[1::3]]
            ms = [p.detach().to(torch.uint8).contiguous() for p in params[2::3]]
            if gemm in SORTED_GEMMS:
                ws, bs, ms = _sort_hidden_units(ws, bs, ms, len(core))
            nbytes = _cabi.flow_blob_bytes(desc, dtype, gemm)
            blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
            _cabi.flow_pack_weights(desc, ws, bs, ms, dtype, gemm, blob)
            # sampling schedule (features that become final per sweep): the words depend on the orders only, so they
            # are built once and copied device-to-device after every re-pack (no host synchronisation per step)
            sched = self._sched_words(device)
            off = _cabi.flow_sched_offset(desc, dtype, gemm)
            blob[off:off + sched.numel()].copy_(sched)
            cached = (key, blob)
            self._blob_cache[gemm] = cached
        base_key = tuple((t.data_ptr(), t._version) for t in self.base.tensors())
        dent = self._desc_cache.get(identity_base)
        if dent is None or dent[0] != base_key:
            dent = (base_key, self.make_desc(identity_base))
            self._desc_cache[identity_base] = dent
        return dent[1], cached[1]


def _sort_hidden_units(ws, bs, ms, n_transforms):
    """Reorder the hidden units of every hyper-network by the number of inputs their mask lets through.

    The order of hidden units is free (rows of W_l, b_l and the matching columns of W_{l+1} move together; the
    function computed is the same up to the summation order inside a dot product).  zuko assigns autoregressive
    degrees round-robin; sorted by degree the masked weight matrices become block-triangular: the warp-level MMA
    kernel stops streaming a 128-column block at its last nonzero weight row (flow_mma_klimit_kernel) and the tcgen05
    kernel issues every k-step only for the accumulator columns that can be nonzero (flow_tc_skip_kernel)."""
    nl = len(ws) // n_transforms
    ws, bs, ms = list(ws), list(bs), list(ms)
    for t in range(n_transforms):
        prev = None
        for l in range(nl):
            i = t * nl + l
            w, b, m = ws[i], bs[i], ms[i]
            if prev is not None:
                w, m = w[:, prev], m[:, prev]
            if l < nl - 1:
                prev = torch.sort(m.sum(dim=1, dtype=torch.int32), stable=True).indices
                w, b, m = w[prev], b[prev], m[prev]
            ws[i], bs[i], ms[i] = w.contiguous(), b.contiguous(), m.contiguous()
    return ws, bs, ms


class NSF(MAF):
    """zuko.flows.NSF: MAF with MonotonicRQSTransform univariates, shapes [(K,), (K,), (K-1,)]."""

    def __init__(self, features, context=0, bins=8, **kwargs):
        kwargs.setdefault("univariate", UNIV_RQS)
        super().__init__(features=features, context=context, bins=bins, **kwargs)


class NCSF(MAF):
    """zuko.flows.NCSF: circular shift + RQS on [-pi, pi], BoxUniform(-pi-1e-5, pi+1e-5) base."""

    def __init__(self, features, context=0, bins=8, **kwargs):
        kwargs.pop("univariate", None)
        kwargs.pop("bound", None)
        super().__init__(features=features, context=context, bins=bins, univariate=UNIV_CIRCULAR_RQS,
                         bound=math.pi, **kwargs)
        self.base = Unconditional(BoxUniform, torch.full((features,), -math.pi - 1e-5),
                                  torch.full((features,), math.pi + 1e-5), buffer=True)
