"""The reference's local flow subclasses, re-expressed on the fused kernels.

  NCSF_gaussian         memflow/transfer_flow/periodicNSF_gaussian.py:24-58
  Custom_spline_flow_ps memflow/unfolding_flow/custom_spline_flow_ps.py:37-81
  Custom_spline_flow_eta memflow/unfolding_flow/custom_spline_flow_eta.py:15-57
  UniformNCSF / UniformNSF / PSNSF   memflow/ttH/transfer_flow/custom_flows.py:58-127
                                     (memflow/models/custom_flows.py has the same classes)
"""
import math

import torch

from .modules import (MAF, BoxUniform, DiagNormal, Unconditional, UNIV_CIRCULAR_RQS, UNIV_PS_RQS,
                      UNIV_RQS)


class NCSF_gaussian(MAF):
    """Circular-shift + RQS(bound=pi) univariates, but MAF's default DiagNormal(0, 1) base."""

    def __init__(self, features, context=0, bins=8, **kwargs):
        super().__init__(features=features, context=context, bins=bins, univariate=UNIV_CIRCULAR_RQS,
                         bound=math.pi, **kwargs)


class Custom_spline_flow_ps(MAF):
    """RQS(bound) univariates with a DiagNormal(mean_gaussian, std_gaussian) base."""

    def __init__(self, features, context=0, bins=8, bound=1.1, mean_gaussian=0.0, std_gaussian=0.35, **kwargs):
        super().__init__(features=features, context=context, bins=bins, univariate=UNIV_RQS, bound=bound,
                         **kwargs)
        self.base = Unconditional(DiagNormal, mean_gaussian * torch.ones((features,)),
                                  std_gaussian * torch.ones((features,)), buffer=True)


class Custom_spline_flow_eta(MAF):
    """RQS(bound=3.5) univariates (``RQSTransform_custom``, custom_spline_flow_eta.py:15-21) on MAF's default
    DiagNormal(0, 1) base -- the pseudorapidity flow."""

    def __init__(self, features, context=0, bins=8, **kwargs):
        kwargs.pop("univariate", None)
        kwargs.pop("bound", None)
        super().__init__(features=features, context=context, bins=bins, univariate=UNIV_RQS, bound=3.5, **kwargs)


class UniformNCSF(MAF):
    def __init__(self, features, context, bins, bound, **kwargs):
        super().__init__(features=features, context=context, bins=bins, univariate=UNIV_CIRCULAR_RQS,
                         bound=bound, **kwargs)
        self.base = Unconditional(BoxUniform, torch.full((features,), -bound - 1e-5),
                                  torch.full((features,), bound + 1e-5), buffer=True)


class UniformNSF(MAF):
    def __init__(self, features, context, bins, bound, **kwargs):
        super().__init__(features=features, context=context, bins=bins, univariate=UNIV_RQS, bound=bound,
                         **kwargs)
        self.base = Unconditional(BoxUniform, torch.full((features,), float(-bound)),
                                  torch.full((features,), float(bound)), buffer=True)


class PSNSF(MAF):
    """PSTransform ([-1,1] -> [0,1], ladj log 1/2) then RQS(bound=1), BoxUniform(0, 1) base."""

    def __init__(self, features, context, bins, **kwargs):
        super().__init__(features=features, context=context, bins=bins, univariate=UNIV_PS_RQS, bound=1.0,
                         **kwargs)
        self.base = Unconditional(BoxUniform, torch.full((features,), 0.0), torch.full((features,), 1.0),
                                  buffer=True)
