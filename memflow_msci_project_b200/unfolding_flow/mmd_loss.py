"""B200 drop-in for ``memflow/unfolding_flow/mmd_loss.py``.

``MMD(x, y, kernel, device, dtype)`` keeps the reference signature (mmd_loss.py:4) and meaning: the empirical
maximum mean discrepancy ``mean(XX + YY - 2 XY)`` of two samples ``x, y [B, D]`` under the sum of rational-quadratic
("multiscale", mmd_loss.py:25-31) or Gaussian ("rbf", mmd_loss.py:33-39) kernels.  The reference builds three B x B Gram
matrices with ``torch.mm`` and one B x B accumulator per bandwidth, and lets autograd differentiate through all of
them; here value AND gradient come from one pass of ``mf_mmd`` (csrc/mmd.cu) that stores nothing of size B^2.
Callers: ``compute_mmd_regr_loss`` / ``compute_mmd_loss`` (scripts/run_model_labframe_sampling.py:114-130).

No CPU fallback: tensors must live on a CUDA device.
"""
import torch

from .. import _cabi

_KERNELS = {"multiscale": _cabi.MF_MMD_MULTISCALE, "rbf": _cabi.MF_MMD_RBF}


class _MMDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, kernel):
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        out = torch.empty(1, dtype=x.dtype, device=x.device)
        gx = torch.empty_like(x) if need_x else None
        gy = torch.empty_like(y) if need_y else None
        _cabi.mmd(x, y, kernel, out, gx, gy)
        ctx.save_for_backward(*(g for g in (gx, gy) if g is not None))
        ctx.have = (need_x, need_y)
        return out.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        saved = list(ctx.saved_tensors)
        gx = saved.pop(0) * grad_out if ctx.have[0] else None
        gy = saved.pop(0) * grad_out if ctx.have[1] else None
        return gx, gy, None


def MMD(x, y, kernel, device=None, dtype=None):
    """Empirical maximum mean discrepancy; the lower, the more evidence that x ~ P and y ~ Q are the same.

    x, y: [B, D] samples (same B, as in the reference where XX, YY and XY share one shape); kernel: "multiscale" or
    "rbf"; ``device`` / ``dtype``: where the reference allocates its accumulators -- the result has ``dtype``
    (default: x.dtype).  Differentiable w.r.t. x and y.  Any other kernel name gives 0 like the reference's
    untouched zero accumulators; B = 0 gives nan (mean of an empty matrix).
    """
    _cabi.require_cuda(x, "x")
    _cabi.require_cuda(y, "y")
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise RuntimeError(f"MMD: x {tuple(x.shape)} and y {tuple(y.shape)} must be [B, D] with the same D")
    if x.shape[0] != y.shape[0]:
        raise RuntimeError(f"MMD: x has {x.shape[0]} rows and y {y.shape[0]}; XX + YY - 2 XY needs the same B")
    dtype = x.dtype if dtype is None else dtype
    if x.shape[0] == 0:
        return torch.full((), float("nan"), dtype=dtype, device=x.device)
    if kernel not in _KERNELS:
        return torch.zeros((), dtype=dtype, device=x.device)
    if x.shape[1] > _cabi.MF_MMD_MAX_DIM:
        raise _cabi.MemflowB200Error(f"MMD: D = {x.shape[1]} > {_cabi.MF_MMD_MAX_DIM} is not supported by mf_mmd")
    return _MMDFn.apply(x.to(dtype).contiguous(), y.to(dtype).contiguous(), _KERNELS[kernel])
