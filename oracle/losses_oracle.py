"""CPU oracle of the periodic-phi regression loss.  TEST INFRASTRUCTURE ONLY: imported by tests/, never
by the product package.

numpy restatement of ``loss_fn_periodic`` (scripts/run_model_labframe_sampling.py:65-77) with loss_fn = HuberLoss /
MSELoss (torch's published definitions), in the arithmetic of the input dtype.  Pinned: tests/test_losses_cpu.py checks
it against tests/golden/losses.npz, which tools/make_golden_losses.py wrote by executing the UNMODIFIED reference
function (extracted from the script's source by ``ast``) with torch's own loss modules.
"""
import numpy as np


def loss_fn_periodic(inp, target, kind="huber", delta=1.0, reduction="mean"):
    inp = np.array(inp, copy=True)
    target = np.asarray(target)
    T = inp.dtype.type
    PI = T(np.pi)
    inp[inp > PI] = inp[inp > PI] - T(2) * PI                     # :68
    inp[inp < -PI] = inp[inp < -PI] + T(2) * PI                   # :69
    dphi = target - inp                                           # :71
    dphi = np.where(dphi > PI, dphi - T(2) * PI, dphi)            # :72
    dphi = np.where(dphi <= -PI, dphi + T(2) * PI, dphi)          # :73
    if kind == "mse":
        l = dphi * dphi
        g = T(2) * dphi
    else:
        ad = np.abs(dphi)
        d = T(delta)
        l = np.where(ad < d, T(0.5) * dphi * dphi, d * (ad - T(0.5) * d))
        g = np.where(ad < d, dphi, np.sign(dphi) * d)
    if reduction == "none":
        return l, g
    total = l.astype(np.float64).sum()
    return (total / l.size if reduction == "mean" else total), g
