"""B200 drop-in for the phase-space entry of ``memflow/unfolding_flow/utils.py``.

``Compute_ParticlesTensor.get_PS`` (utils.py:1056-1152) is on the hot path; it is the n = 4
RAMBO inverse without Jacobian plus the boolean ``mask_problems``.  ``get_PS_fromlab`` is the fused form of the
chain in front of it in the model wrapper (boost fix-up, lab -> CM boost, get_PS;
memflow/unfolding_flow/unfolding_flow.py:137-163) as ONE launch.  The small coordinate helpers the training
scripts call on kernel outputs (``get_ptetaphi_comp``, ``get_cartesian_comp``; utils.py:109-185) are kept as
differentiable torch expressions.  The reference defines the
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

    @staticmethod
    def get_PS_fromlab(data_regressed_lab, boost_regressed, order=[0, 1, 2, 3],
                       target_mass=torch.Tensor([[M_HIGGS, M_TOP, M_TOP, M_GLUON]]), E_CM=13000,
                       fix_boost=True, compute="f64"):
        """Fused memflow/unfolding_flow/unfolding_flow.py:137-163: lab-frame momenta [N,4,4] + regressed boost [N,4]
        -> (data_regressed_cm [N,4,4], boost_fixed [N,4], ps [N,10], 0*jac [N], mask_problems [N]).

        One launch of ``mf_rambo_get_ps_lab``: boost fix-up (``fix_boost``: energies with sqrt(E^2 - pz^2) below
        M_MIN_TOT are raised like :140-143), ``boost_tt(lab, -boostVector_t(boost))``, then get_PS."""
        _cabi.require_cuda(data_regressed_lab, "data_regressed_lab")
        perm = [int(i) for i in order]
        key = (tuple(perm), tuple(float(v) for v in target_mass.reshape(-1)), float(E_CM))
        desc = _desc_cache.get(key)
        if desc is None:
            desc = build_rambo_desc(target_mass[:, perm].reshape(-1), E_CM, sum_mass=target_mass.sum())
            _desc_cache[key] = desc
        P = data_regressed_lab
        if P.dtype not in (torch.float32, torch.float64):
            P = P.float()
        P = P.contiguous()
        boost = boost_regressed.to(P.dtype).contiguous()
        N = P.shape[0]
        cm = torch.empty_like(P)
        boost_fixed = torch.empty_like(boost)
        ps = torch.empty((N, 10), dtype=P.dtype, device=P.device)
        mask = torch.empty(N, dtype=torch.bool, device=P.device)
        cmode = _cabi.MF_COMPUTE_F64 if (compute == "f64" or P.dtype == torch.float64) else _cabi.MF_COMPUTE_F32
        _cabi.rambo_get_ps_lab(desc, P, boost, None if perm == [0, 1, 2, 3] else perm,
                               Compute_ParticlesTensor.M_MIN_TOT if fix_boost else 0.0, cm, boost_fixed, ps, mask, cmode)
        return cm, boost_fixed, ps, torch.zeros(N, dtype=P.dtype, device=P.device), mask

    # ---- coordinate helpers (utils.py:109-185), torch expressions on any device, differentiable ----
    @staticmethod
    def get_cartesian_comp(particle, mass):
        """(pt, eta, phi) [N,3] + mass -> (E, px, py, pz) [N,4] (utils.py:109-119)."""
        Px = particle[:, 0] * torch.cos(particle[:, 2])
        Py = particle[:, 0] * torch.sin(particle[:, 2])
        Pz = particle[:, 0] * torch.sinh(particle[:, 1])
        E = torch.sqrt(Px ** 2 + Py ** 2 + Pz ** 2 + mass ** 2)
        return torch.stack((E, Px, Py, Pz), dim=1)

    @staticmethod
    def get_ptetaphi_comp(particle):
        """(E, px, py, pz) [N,4] -> (E, pt, eta, phi) [N,4] (utils.py:169-173)."""
        pt = (particle[:, 1] ** 2 + particle[:, 2] ** 2) ** 0.5
        eta = -torch.log(torch.tan(torch.atan2(pt, particle[:, 3]) / 2))
        phi = torch.atan2(particle[:, 2], particle[:, 1])
        return torch.stack((particle[:, 0], pt, eta, phi), dim=1)

    @staticmethod
    def get_ptetaphi_comp_batch(particles):
        """[N,P,4] version (utils.py:175-179)."""
        pt = (particles[..., 1] ** 2 + particles[..., 2] ** 2) ** 0.5
        eta = -torch.log(torch.tan(torch.atan2(pt, particles[..., 3]) / 2))
        phi = torch.atan2(particles[..., 2], particles[..., 1])
        return torch.stack((particles[..., 0], pt, eta, phi), dim=2)
