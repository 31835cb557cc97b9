"""zuko-shaped conditional spline flows on fused sm_100a kernels (see modules.py)."""
from .modules import (MAF, NCSF, NSF, BoxUniform, DiagNormal, MaskedAutoregressiveTransform, MaskedLinear,
                      MaskedMLP, NormalizingFlow, SimpleAffineTransform, Unconditional,
                      UnconditionalDistribution, SigmoidAffine, LogitAffine)
from .transforms import MonotonicRQSTransform
from .custom import NCSF_gaussian, Custom_spline_flow_ps, Custom_spline_flow_eta, PSNSF, UniformNCSF, UniformNSF

__all__ = ["MAF", "NSF", "NCSF", "BoxUniform", "DiagNormal", "MaskedAutoregressiveTransform", "MaskedLinear",
           "MaskedMLP", "NormalizingFlow", "SimpleAffineTransform", "Unconditional",
           "UnconditionalDistribution", "NCSF_gaussian", "Custom_spline_flow_ps", "PSNSF", "UniformNCSF",
           "UniformNSF", "MonotonicRQSTransform", "Custom_spline_flow_eta", "SigmoidAffine", "LogitAffine"]
