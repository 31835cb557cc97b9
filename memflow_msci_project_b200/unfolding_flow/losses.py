"""B200 drop-in for the periodic-phi regression loss of the reference's training scripts.

``loss_fn_periodic(inp, target, loss_fn, device)`` keeps the signature and meaning of
scripts/run_model_labframe_sampling.py:65-77 (the same function is re-defined in every ``run_model_labframe*`` script):
``inp`` wrapped once into [-pi, pi], ``target - inp`` wrapped into (-pi, pi], then ``loss_fn(delta_phi, 0)``.  The
reference spends about ten elementwise torch kernels on it, four times per training step
(compute_regr_losses :100-111); here ``loss_fn`` must be a ``torch.nn.HuberLoss`` or ``torch.nn.MSELoss`` instance (the
only ones the scripts use) and the whole chain, the reduction and the derivative for the backward are ``mf_periodic_loss``
(csrc/periodic.cu).  Unlike the reference, ``inp`` is not modified in place (the scripts pass a temporary).

No CPU fallback: tensors must live on a CUDA device.
"""
import torch

from .. import _cabi


class _PeriodicLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, target, kind, delta, reduction):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        deriv = torch.empty_like(inp) if need else None
        if reduction == "none":
            out = torch.empty_like(inp)
            _cabi.periodic_loss(inp, target, kind, delta, False, elem=out, deriv=deriv)
        else:
            out = torch.empty(1, dtype=inp.dtype, device=inp.device)
            _cabi.periodic_loss(inp, target, kind, delta, reduction == "mean", deriv=deriv, total=out)
            out = out.reshape(())
        if need:
            ctx.save_for_backward(deriv)
        ctx.scale = 1.0 / inp.numel() if reduction == "mean" else 1.0
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (deriv,) = ctx.saved_tensors
        g = deriv * (grad_out * ctx.scale)                 # d loss / d delta_phi; d/d target = +, d/d inp = -
        return (-g if ctx.needs_input_grad[0] else None, g if ctx.needs_input_grad[1] else None, None, None, None)


def loss_fn_periodic(inp, target, loss_fn, device=None):
    """Loss of a periodic variable: ``loss_fn(wrap(target - wrap(inp)), 0)``; differentiable w.r.t. inp and target."""
    _cabi.require_cuda(inp, "inp")
    _cabi.require_cuda(target, "target")
    if isinstance(loss_fn, torch.nn.HuberLoss):
        kind, delta = _cabi.MF_LOSS_HUBER, float(loss_fn.delta)
    elif isinstance(loss_fn, torch.nn.MSELoss):
        kind, delta = _cabi.MF_LOSS_MSE, 0.0
    else:
        raise TypeError(f"loss_fn_periodic: {type(loss_fn).__name__} is not fused on the B200 path "
                        "(torch.nn.HuberLoss and torch.nn.MSELoss are)")
    if loss_fn.reduction not in ("mean", "sum", "none"):
        raise ValueError(f"loss_fn_periodic: reduction {loss_fn.reduction!r}")
    if inp.shape != target.shape:
        raise RuntimeError(f"loss_fn_periodic: inp {tuple(inp.shape)} and target {tuple(target.shape)} differ")
    if inp.dtype != target.dtype:
        raise RuntimeError(f"loss_fn_periodic: inp is {inp.dtype} and target {target.dtype}")
    if inp.numel() == 0:
        return loss_fn(inp - target, torch.zeros_like(inp))   # nan (mean) / 0 (sum) / empty, as torch defines them
    return _PeriodicLossFn.apply(inp.contiguous(), target.contiguous(), kind, delta, loss_fn.reduction)
