"""Autograd bridges of the fused flow kernels (training path, BASELINE config 4).

STATUS (round 2): the FORWARD of a differentiable ``log_prob`` / ``rsample`` is the fused sm_100a kernel;
the BACKWARD re-evaluates the flow transform by transform on the GPU: the hyper-network (masked linears + ReLU)
through `_MaskedLinearFn`, whose three GEMMs per layer (recompute, dX, dW) run on tcgen05 in fp32-class 3xFP16
arithmetic (`mf_gemm_nt_3xf16`; fp32 only -- fp64 training keeps torch's GEMMs), the spline with native kernels --
forward `mf_rqs_transform`, backward `mf_rqs_backward` (closed-form gradient w.r.t. the 3K-1 logits and x; the
inverse direction by the implicit function theorem) -- instead of ~300 eager elementwise kernels per transform.  This keeps the reference's
training loops working unchanged (`loss = -flow(c).log_prob(x).mean(); loss.backward()`); it never runs on the CPU
and never touches oracle/.  `MEMFLOW_B200_EAGER_SPLINE=1` selects the pure-torch spline restated below (the
formulation the native backward is tested against).
"""
import os
import math

import torch
import torch.nn.functional as F

from .. import _cabi
from .modules import BASE_DIAG_NORMAL, UNIV_CIRCULAR_RQS, UNIV_PS_RQS


def _softclip(v, factor):
    return v / (1 + torch.abs(factor * v))


def _rqs_knots(w, h, d, bound, slope):
    if slope:
        ls = math.log(slope)
        w, h, d = _softclip(w, 2 / ls), _softclip(h, 2 / ls), _softclip(d, 1 / ls)
    W = F.pad(F.softmax(w, dim=-1), (1, 0), value=0)
    H = F.pad(F.softmax(h, dim=-1), (1, 0), value=0)
    D = F.pad(d, (1, 1), value=0)
    return bound * (2 * torch.cumsum(W, -1) - 1), bound * (2 * torch.cumsum(H, -1) - 1), torch.exp(D)


def _bin(k, hor, ver, der):
    K = hor.shape[-1] - 1
    mask = (0 <= k) & (k < K)
    k = k % K
    idx = torch.stack((k, k + 1), dim=-1)
    x0, x1 = hor.gather(-1, idx).unbind(-1)
    y0, y1 = ver.gather(-1, idx).unbind(-1)
    d0, d1 = der.gather(-1, idx).unbind(-1)
    return mask, x0, x1, y0, y1, d0, d1, (y1 - y0) / (x1 - x0)


def _rqs_forward(x, hor, ver, der):
    k = (hor < x[..., None]).sum(-1) - 1
    mask, x0, x1, y0, y1, d0, d1, s = _bin(k, hor, ver, der)
    z = mask * (x - x0) / (x1 - x0)
    den = s + (d0 + d1 - 2 * s) * z * (1 - z)
    y = y0 + (y1 - y0) * (s * z**2 + d0 * z * (1 - z)) / den
    jac = s**2 * (2 * s * z * (1 - z) + d0 * (1 - z) ** 2 + d1 * z**2) / den**2
    return torch.where(mask, y, x), mask * torch.log(torch.where(mask, jac, torch.ones_like(jac)))


def _rqs_inverse(y, hor, ver, der):
    k = (ver < y[..., None]).sum(-1) - 1
    mask, x0, x1, y0, y1, d0, d1, s = _bin(k, hor, ver, der)
    y_ = mask * (y - y0)
    a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - 2 * s)
    b = (y1 - y0) * d0 - y_ * (d0 + d1 - 2 * s)
    c = -s * y_
    z = 2 * c / (-b - torch.sqrt(b**2 - 4 * a * c))
    return torch.where(mask, x0 + z * (x1 - x0), y)


def _wrap(x, bound):
    return torch.remainder(x + bound, 2 * bound) - bound


def _native_gemm():
    """fp32 hyper-network GEMMs of the recompute / backward on tcgen05 (mf_gemm_nt_3xf16) instead of cuBLAS;
    MEMFLOW_B200_TORCH_GEMM=1 keeps torch's (the formulation the native path is tested against)."""
    return os.environ.get("MEMFLOW_B200_TORCH_GEMM") != "1"


def _pow2_scale(t):
    """device scalar 2^s that lifts the largest magnitude of t to ~2^14 (fp16 tops out at 2^16; the lo part of a value stays a normal fp16 number down to 2^-17 of the largest one) (no host sync): gradients are far below the
    fp16 range of the split operands of mf_gemm_nt_3xf16"""
    amax = torch.linalg.vector_norm(t, ord=float("inf")).clamp_min(1e-30)
    return torch.exp2(torch.floor(14.0 - torch.log2(amax))).reshape(1).to(torch.float32)


class _MaskedLinearFn(torch.autograd.Function):
    """y = [relu](x @ (mask * weight).T + bias) with all three GEMMs (forward, dX, dW) on tcgen05 in fp32-class 3xFP16
    arithmetic: the native hyper-network backward."""

    @staticmethod
    def forward(ctx, x, weight, mask, bias, relu, chained=False, input_is_relu=False):
        """chained: this layer sits inside the masked MLP, i.e. its input is the ReLU output of the layer below (unless
        it is the first layer) and its output feeds the next `_MaskedLinearFn`.  Then the input gradient leaves this
        layer already multiplied by the ReLU derivative of the layer below (x > 0, applied in the GEMM epilogue), and
        the gradient this layer receives arrives masked the same way by the layer above -- no mask operand in the
        staging of either GEMM, and the weight gradient can take its operands by bulk copy."""
        ctx.chained = chained
        ctx.input_is_relu = input_is_relu
        wm = (mask * weight).contiguous()
        x = x.contiguous()
        y = torch.empty((x.shape[0], wm.shape[0]), dtype=x.dtype, device=x.device)
        _cabi.gemm_nt(x, wm, y, bias=bias.contiguous(), relu=relu)
        ctx.relu = relu
        ctx.save_for_backward(x, wm, mask, y if relu else x.new_empty(0))
        return y

    @staticmethod
    def backward(ctx, gy):
        x, wm, mask, y = ctx.saved_tensors
        gy = gy.contiguous()
        # the ReLU derivative (y > 0): inside the GEMMs' operand staging, unless the layer above already applied it
        relu_of = y if (ctx.relu and not ctx.chained) else None
        scale = _pow2_scale(gy)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            # chained and fed by a ReLU (every layer but the first): gx leaves masked with (x > 0) for the layer below
            fed_by_relu = ctx.chained and ctx.input_is_relu
            _cabi.gemm_nt(gy, wm.t().contiguous(), gx, mask=relu_of, out_mask=x if fed_by_relu else None, scale_a=scale)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[3]:
            # dW = dZ^T X and db = column sums of dZ in ONE launch straight from the row-major activations; K = rows,
            # cut into pieces whose partial sums are added into gw / gb
            out, inn = wm.shape
            gw = torch.zeros_like(wm)
            gb = torch.zeros(out, dtype=wm.dtype, device=wm.device)
            k_split = max(1, min(592, x.shape[0] // 512))
            if 128 < out <= 256 and inn <= 128 and relu_of is None:
                # a few rows over one 128-row tile (the last layer, 141 outputs): the transposed product X^T dY has ONE
                # row tile instead of a second, nearly empty one that would stage the whole of X again
                gwt = torch.zeros((inn, out), dtype=wm.dtype, device=wm.device)
                _cabi.gemm_tn(x, gy, gwt, k_split=k_split, scale_b=scale)
                gw, gb = gwt.t(), gy.sum(0)
            elif inn + 1 <= 256:
                _cabi.gemm_tn(gy, x, gw, k_split=k_split, mask=relu_of, colsum=gb, scale_a=scale)
            else:  # wider than one accumulator tile: torch
                gz = gy * (y > 0) if ctx.relu else gy
                gw, gb = gz.t() @ x, gz.sum(0)
            gw = gw * mask
        return gx, gw, None, gb, None, None, None


def _linear(h, lin, relu, first=True, chained=False):
    if _native_gemm() and h.is_cuda and h.dtype == torch.float32 and h.dim() == 2:
        return _MaskedLinearFn.apply(h, lin.weight, lin.mask, lin.bias, relu, chained, not first)
    h = F.linear(h, lin.mask * lin.weight, lin.bias)
    return torch.relu(h) if relu else h


def _hyper_phi(t, x, c):
    h = x if c is None else torch.cat((x, c), dim=-1)
    lins = t.hyper.linears()
    for i, lin in enumerate(lins):
        h = _linear(h, lin, i < len(lins) - 1, first=(i == 0), chained=True)
    return h.unflatten(-1, (t.features, t.total))


def _hyper(t, x, c):
    phi = _hyper_phi(t, x, c)
    K = t.bins
    return _rqs_knots(phi[..., :K], phi[..., K:2 * K], phi[..., 2 * K:], t.bound, t.slope)


def _native_spline():
    return os.environ.get("MEMFLOW_B200_EAGER_SPLINE") != "1"


class _SplineForwardFn(torch.autograd.Function):
    """(phi [n, 3K-1], x [n]) -> (y, ladj): mf_rqs_transform forward, mf_rqs_backward backward."""

    @staticmethod
    def forward(ctx, phi, x, K, bound, slope):
        y, ladj = torch.empty_like(x), torch.empty_like(x)
        _cabi.rqs_transform(phi, x, K, 0, bound, slope, False, y, ladj)
        ctx.save_for_backward(phi, x)
        ctx.meta = (K, bound, slope)
        return y, ladj

    @staticmethod
    def backward(ctx, gy, gl):
        phi, x = ctx.saved_tensors
        K, bound, slope = ctx.meta
        gphi, gx = torch.empty_like(phi), torch.empty_like(x)
        _cabi.rqs_backward(phi, x, K, bound, slope, gy.contiguous(), gl.contiguous(), gphi, gx)
        return gphi, gx, None, None, None


class _SplineInverseFn(torch.autograd.Function):
    """(phi, y) -> x = F^-1(y; phi).  With F(x, phi) = y: dx = (dy - dF/dphi dphi) / F_x, F_x = exp(ladj(x))."""

    @staticmethod
    def forward(ctx, phi, y, K, bound, slope):
        x = torch.empty_like(y)
        _cabi.rqs_transform(phi, y, K, 0, bound, slope, True, x, None)
        ctx.save_for_backward(phi, x)
        ctx.meta = (K, bound, slope)
        return x

    @staticmethod
    def backward(ctx, gx):
        phi, x = ctx.saved_tensors
        K, bound, slope = ctx.meta
        y2, ladj = torch.empty_like(x), torch.empty_like(x)
        _cabi.rqs_transform(phi, x, K, 0, bound, slope, False, y2, ladj)
        gy = gx * torch.exp(-ladj)
        gphi, scratch = torch.empty_like(phi), torch.empty_like(x)
        _cabi.rqs_backward(phi, x, K, bound, slope, (-gy).contiguous(), None, gphi, scratch)
        return gphi, gy, None, None, None


def _spline_forward(t, phi, xin):
    """Spline of transform t on every (row, feature): phi [N, D, 3K-1], xin [N, D] -> (y, ladj) [N, D]."""
    shape = xin.shape
    y, l = _SplineForwardFn.apply(phi.reshape(-1, t.total).contiguous(), xin.reshape(-1).contiguous(), t.bins,
                                  float(t.bound), float(t.slope or 0.0))
    return y.view(shape), l.view(shape)


def _spline_inverse(t, phi, y):
    shape = y.shape
    x = _SplineInverseFn.apply(phi.reshape(-1, t.total).contiguous(), y.reshape(-1).contiguous(), t.bins,
                               float(t.bound), float(t.slope or 0.0))
    return x.view(shape)


def eager_forward(module, x, c):
    """Differentiable restatement: (z, ladj) of the spline transforms of ``module`` (rows x [N,D], c [N,C])."""
    ladj = x.new_zeros(x.shape[:-1])
    native = _native_spline() and x.is_cuda
    for t in module.core_transforms():
        xin, l0 = x, 0.0
        if t.univariate == UNIV_CIRCULAR_RQS:
            xin = _wrap(x, t.bound)
        elif t.univariate == UNIV_PS_RQS:
            xin, l0 = (x + 1) / 2, math.log(0.5)
        if native:
            x, l = _spline_forward(t, _hyper_phi(t, x, c), xin)
        else:
            x, l = _rqs_forward(xin, *_hyper(t, x, c))
        ladj = ladj + (l + l0).sum(-1)
    return x, ladj


def eager_inverse(module, z, c):
    y = z
    native = _native_spline() and z.is_cuda
    for t in reversed(module.core_transforms()):
        x = torch.zeros_like(y)
        for _ in range(t.passes):
            if native:
                x = _spline_inverse(t, _hyper_phi(t, x, c), y)
            else:
                x = _rqs_inverse(y, *_hyper(t, x, c))
            if t.univariate == UNIV_CIRCULAR_RQS:
                x = _wrap(x, t.bound)
            elif t.univariate == UNIV_PS_RQS:
                x = 2 * x - 1
        y = x
    return y


def _params(module):
    ps = []
    for t in module.core_transforms():
        for lin in t.hyper.linears():
            ps += [lin.weight, lin.bias]
    return ps


class _FlowForwardFn(torch.autograd.Function):
    """(x, c, *params) -> (z, ladj): fused kernel forward, eager-torch recompute backward."""

    @staticmethod
    def forward(ctx, bound_flow, x, c, *params):
        m = bound_flow._m
        dtype = m.compute_dtype()
        G = bound_flow._ctx_group
        gemm = m.effective_gemm(dtype, ctx_group=G)
        desc, blob = m.packed(x.device, dtype, gemm=gemm)
        desc = bound_flow._with_group(desc)   # ctx_group > 1: c has one row per G consecutive x rows
        z = torch.empty_like(x)
        ladj = torch.empty(x.shape[0], dtype=dtype, device=x.device)
        _cabi.flow_log_prob(desc, blob, x, c, None, z, ladj, None, None, gemm)
        ctx.module = m
        ctx.group = G
        ctx.has_c = c is not None
        ctx.save_for_backward(x, c if c is not None else x.new_empty(0))
        return z, ladj

    @staticmethod
    def backward(ctx, gz, gladj):
        m = ctx.module
        x, c = ctx.saved_tensors
        c = c if ctx.has_c else None
        params = _params(m)
        with torch.enable_grad():
            xr = x.detach().requires_grad_(True)
            cr = c.detach().requires_grad_(True) if c is not None else None
            # grouped context: the recompute sees the context repeated per row; autograd sums the group's gradient
            ce = cr.repeat_interleave(ctx.group, dim=0) if (cr is not None and ctx.group > 1) else cr
            z, ladj = eager_forward(m, xr, ce)
            inputs = [xr] + ([cr] if cr is not None else []) + [p for p in params if p.requires_grad]
            grads = torch.autograd.grad([z, ladj], inputs, [gz, gladj], allow_unused=True)
        it = iter(grads)
        gx = next(it)
        gc = next(it) if cr is not None else None
        gp = [next(it) if p.requires_grad else None for p in params]
        return (None, gx, gc, *gp)


def flow_log_prob_autograd(bound_flow, x, want_logprob):
    """Differentiable counterpart of NormalizingFlow._forward (same return triple)."""
    m = bound_flow._m
    pre, core, post = bound_flow._split()
    if pre or post:
        raise NotImplementedError("gradients through SimpleAffineTransform wrappers are not implemented")
    _cabi.require_cuda(x, "x")
    dtype = m.compute_dtype()
    xb, cb, batch = bound_flow._rows(x.to(dtype))
    xb = xb.contiguous()
    cb = cb.to(dtype).contiguous() if cb is not None else None
    z, ladj = _FlowForwardFn.apply(bound_flow, xb, cb, *_params(m))
    logprob = None
    if want_logprob:
        base = m.base()
        logprob = (base.log_prob(z) + ladj).reshape(batch)
    return z.reshape(*batch, m.features), ladj.reshape(batch), logprob


class _FlowInverseFn(torch.autograd.Function):
    """(z, c, *params) -> x = T^-1(z): fused sampling kernel forward, eager-torch recompute backward."""

    @staticmethod
    def forward(ctx, bound_flow, z, c, *params):
        m = bound_flow._m
        dtype = m.compute_dtype()
        G = bound_flow._ctx_group
        gemm = m.effective_gemm(dtype, sampling=True, ctx_group=G)
        desc, blob = m.packed(z.device, dtype, identity_base=True, gemm=gemm)
        desc = bound_flow._with_group(desc)
        x = torch.empty_like(z)
        _cabi.flow_sample(desc, blob, z, c, x, gemm)
        ctx.module = m
        ctx.group = G
        ctx.has_c = c is not None
        ctx.save_for_backward(z, c if c is not None else z.new_empty(0))
        return x

    @staticmethod
    def backward(ctx, gx):
        m = ctx.module
        z, c = ctx.saved_tensors
        c = c if ctx.has_c else None
        params = _params(m)
        with torch.enable_grad():
            zr = z.detach().requires_grad_(True)
            cr = c.detach().requires_grad_(True) if c is not None else None
            ce = cr.repeat_interleave(ctx.group, dim=0) if (cr is not None and ctx.group > 1) else cr
            x = eager_inverse(m, zr, ce)
            inputs = [zr] + ([cr] if cr is not None else []) + [p for p in params if p.requires_grad]
            grads = torch.autograd.grad([x], inputs, [gx], allow_unused=True)
        it = iter(grads)
        gz = next(it)
        gc = next(it) if cr is not None else None
        gp = [next(it) if p.requires_grad else None for p in params]
        return (None, gz, gc, *gp)


def flow_rsample_autograd(bound_flow, z):
    m = bound_flow._m
    pre, core, post = bound_flow._split()
    if pre or post:
        raise NotImplementedError("gradients through SimpleAffineTransform wrappers are not implemented")
    dtype = m.compute_dtype()
    zb, cb, batch = bound_flow._rows(z.to(dtype))
    zb = zb.contiguous()
    cb = cb.to(dtype).contiguous() if cb is not None else None
    x = _FlowInverseFn.apply(bound_flow, zb, cb, *_params(m))
    return x.reshape(*batch, m.features)
