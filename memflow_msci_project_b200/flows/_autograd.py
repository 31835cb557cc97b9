"""Autograd bridges of the fused flow kernels (training path, BASELINE config 4)."""
from .. import _cabi


def flow_log_prob_autograd(bound_flow, x, want_logprob):
    raise NotImplementedError(
        "gradients through the fused flow kernel are not implemented yet: call under torch.no_grad() "
        "(a silent fallback to eager torch is deliberately not provided)")


def flow_rsample_autograd(bound_flow, z):
    raise NotImplementedError(
        "gradients through the fused sampling kernel are not implemented yet: call under torch.no_grad()")
