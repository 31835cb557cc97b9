"""Standalone univariate transforms with zuko's names, on the fused spline kernel (mf_rqs_transform).

``MonotonicRQSTransform(widths, heights, derivatives, bound=5.0, slope=1e-3)`` is what the reference's
local univariates wrap (memflow/ttH/transfer_flow/custom_flows.py:33-56,
memflow/unfolding_flow/custom_spline_flow_ps.py:16-25); the flows themselves never materialise the
parameters, this class serves callers that already hold them.
"""
import torch

from .. import _cabi
from .modules import UNIV_CIRCULAR_RQS, UNIV_PS_RQS, UNIV_RQS  # noqa: F401


class MonotonicRQSTransform:
    """widths/heights [..., K], derivatives [..., K-1] (unconstrained), acting element-wise on x [...]."""

    def __init__(self, widths, heights, derivatives, bound=5.0, slope=1e-3, univariate=UNIV_RQS):
        K = widths.shape[-1]
        if heights.shape[-1] != K or derivatives.shape[-1] != K - 1:
            raise ValueError("expected widths [...,K], heights [...,K], derivatives [...,K-1]")
        _cabi.require_cuda(widths, "widths")
        self.bins, self.bound, self.slope, self.univariate = K, float(bound), slope, univariate
        self.params = torch.cat((widths, heights, derivatives), dim=-1).contiguous()  # [..., 3K-1]

    @classmethod
    def from_packed(cls, phi, bins, bound=5.0, slope=1e-3, univariate=UNIV_RQS):
        """phi [..., 3K-1] exactly as the hyper-network emits it (no copy if contiguous)."""
        self = cls.__new__(cls)
        _cabi.require_cuda(phi, "phi")
        if phi.shape[-1] != 3 * bins - 1:
            raise ValueError(f"expected last dimension {3 * bins - 1}, got {tuple(phi.shape)}")
        self.bins, self.bound, self.slope, self.univariate = bins, float(bound), slope, univariate
        self.params = phi.contiguous()
        return self

    def _run(self, x, inverse, debug=False):
        p = self.params
        x = x.to(p.dtype).expand(p.shape[:-1]).contiguous()
        y = torch.empty_like(x)
        ladj = None if inverse else torch.empty_like(x)
        bins = torch.empty(x.shape, dtype=torch.int8, device=x.device) if debug else None
        inside = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if debug else None
        _cabi.rqs_transform(p, x, self.bins, self.univariate, self.bound, self.slope, inverse, y, ladj, bins, inside)
        return y, ladj, bins, (inside.bool() if debug else None)

    def __call__(self, x):
        return self._run(x, False)[0]

    def call_and_ladj(self, x):
        y, ladj, _, _ = self._run(x, False)
        return y, ladj

    def log_abs_det_jacobian(self, x, y=None):
        return self._run(x, False)[1]

    def inv(self, y):
        return self._run(y, True)[0]

    def call_debug(self, x):
        return self._run(x, False, debug=True)
