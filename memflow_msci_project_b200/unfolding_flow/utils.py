"""B200 drop-in for the phase-space entry of ``memflow/unfolding_flow/utils.py``.

Only ``Compute_ParticlesTensor.get_PS`` (utils.py:1056-1152) is on the hot path; it is the n = 4
RAMBO inverse without Jacobian plus the boolean ``mask_problems``.  The reference defines the
methods of this class without ``self`` and calls them on the class itself
(``particle_tools.get_PS(...)``, memflow/unfolding_flow/unfolding_flow.py:163), hence staticmethods.
"""
import torch

from .. import _cabi
from ..phasespace._desc import build_rambo_desc

M_HIGGS = 125.25
M_TOP = 172.5
M_GLUON = 1e-5

_desc_cache = {}


class Compute_ParticlesTensor:

    M_MIN_TOT = M_HIGGS + 2 * M_TOP + M_GLUON

    @staticmethod
    def get_PS(data_HttISR, boost_reco, order=[0, 1, 2, 3],
               target_mass=torch.Tensor([[M_HIGGS, M_TOP, M_TOP, M_GLUON]]), E_CM=13000,
               compute="f64"):
        """momenta [N,4,4] (CM frame) + boost [N,4] -> (ps [N,10], 0*jac [N], mask_problems [N]).

        One launch of ``mf_rambo_get_ps``; never raises on bad kinematics, the mask reports them.
        """
        _cabi.require_cuda(data_HttISR, "data_HttISR")
        perm = [int(i) for i in order]
        key = (tuple(perm), tuple(float(v) for v in target_mass.reshape(-1)), float(E_CM))
        desc = _desc_cache.get(key)
        if desc is None:
            desc = build_rambo_desc(target_mass[:, perm].reshape(-1), E_CM, sum_mass=target_mass.sum())
            _desc_cache[key] = desc
        P = data_HttISR
        if P.dtype not in (torch.float32, torch.float64):
            P = P.float()
        P = P.contiguous()
        boost = boost_reco.to(P.dtype).contiguous()
        N = P.shape[0]
        ps = torch.empty((N, 10), dtype=P.dtype, device=P.device)
        mask = torch.empty(N, dtype=torch.bool, device=P.device)
        cmode = _cabi.MF_COMPUTE_F64 if (compute == "f64" or P.dtype == torch.float64) else _cabi.MF_COMPUTE_F32
        _cabi.rambo_get_ps(desc, P, boost, None if perm == [0, 1, 2, 3] else perm, ps, mask, cmode)
        # detjinv_regressed = 0 in the reference (:1139): the returned "jacobian" is identically 0
        return ps, torch.zeros(N, dtype=P.dtype, device=P.device), mask
