"""Four-vector helpers with the names of ``memflow/phasespace/utils.py``.

The model wrappers of the reference call a handful of these directly (``boostVector_t``,
``boost_tt`` at memflow/unfolding_flow/unfolding_flow.py:147-148), so the drop-in keeps them as
plain torch expressions (device-agnostic glue; the fused kernels do not use them).  Tensors are
``[..., 4]`` with components (E, px, py, pz).
"""
import math

import torch


def rho2_t(p):
    """|p_vec|^2 of ``[N,4]`` vectors (phasespace/utils.py:25-28)."""
    return (p[:, 1:] * p[:, 1:]).sum(-1)


def rho2_tt(p):
    """|p_vec|^2 of ``[N,P,4]`` vectors (:31-34)."""
    return (p[:, :, 1:] ** 2).sum(-1)


def dot_t(a, b):
    """Minkowski product of ``[N,4]`` vectors (:56-65)."""
    return a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2] - a[:, 3] * b[:, 3]


def dot_s(a, b):
    return a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]


def dot_fb(a, b):
    return a[:, :, 0] * b[:, :, 0] + a[:, :, 1] * b[:, :, 1] + a[:, :, 2] * b[:, :, 2]


def square_t(p):
    """Minkowski square for 4-component rows, Euclidean square otherwise (:45-51).  Unlike the
    reference the dispatch looks at the LAST dimension only, so a batch of 4 three-vectors is
    handled correctly."""
    if p.shape[-1] == 4:
        return p[..., 0] ** 2 - (p[..., 1:] ** 2).sum(-1)
    return (p * p).sum(-1)


def mass_t(p):
    return torch.sqrt(square_t(p))


def set_square_t(p, square, negative=False):
    """Copy of ``p`` whose energy makes its Minkowski square equal ``square`` (:6-22)."""
    out = p.clone()
    out[:, 0] = (rho2_t(p) + square) ** 0.5
    if negative:
        out[:, 0] = -out[:, 0]
    return out


def boostVector_t(p):
    """beta = p_vec / E (:36-42); no host-syncing validity print."""
    return p[:, 1:] / p[:, 0].unsqueeze(1)


def _boost(p, beta, gamma):
    b2 = (beta * beta).sum(-1)
    if not torch.is_tensor(gamma) and gamma < 0.0:
        gamma = 1.0 / torch.sqrt(1.0 - b2)
    bp = (p[..., 1:] * beta).sum(-1)
    gamma2 = torch.where(b2 > 0, (gamma - 1.0) / b2, torch.zeros_like(b2))
    factor = gamma2 * bp + gamma * p[..., 0]
    space = p[..., 1:] + factor.unsqueeze(-1) * beta
    energy = gamma * (p[..., 0] + bp)
    return torch.cat((energy.unsqueeze(-1), space), dim=-1)


def boost_t(p, boost_vector, gamma=-1.0):
    """Boost ``[N,4]`` by ``[N,3]`` velocities (:88-118): boost_t(p, -boostVector_t(p)) = (M,0,0,0)."""
    return _boost(p, boost_vector, gamma)


def boost_tt(p, boost_vector, gamma=-1.0):
    """Boost ``[N,P,4]`` by ``[N,1,3]`` (or ``[N,P,3]``) velocities (:121-146)."""
    return _boost(p, boost_vector.expand(p.shape[0], p.shape[1], 3), gamma)


def boost_to_lab_frame(momenta, xb_1, xb_2):
    """CM -> lab boost from the Bjorken fractions (:180-202)."""
    ref = momenta[:, 0, :] * xb_1.unsqueeze(-1) + momenta[:, 1, :] * xb_2.unsqueeze(-1)
    moving = (rho2_t(ref) != 0) & ((xb_1 != 1) | (xb_2 != 1))
    beta = torch.where(moving.unsqueeze(-1), boostVector_t(ref), torch.zeros_like(ref[:, 1:]))
    return boost_tt(momenta, beta.unsqueeze(1))


def pseudoRap(p, eps=2.220446049250313e-16 ** 0.5, huge=1.7976931348623157e308):
    pt = torch.sqrt((p[..., 1:3] ** 2).sum(-1))
    th = torch.atan2(pt, p[..., 3])
    return torch.where((pt < eps) & (p[..., 3].abs() < eps), torch.full_like(pt, huge),
                       -torch.log(torch.tan(th / 2.0)))


pseudoRap_t = pseudoRap


def uniform_distr(r, minv, maxv):
    d = torch.ones_like(r) * maxv - torch.ones_like(r) * minv
    return torch.ones_like(r) * minv + d * r, d


def get_x1x2_from_uniform(unif, final_state_mass, E_cm):
    """Uniform pair -> (x1, x2, weight) (:260-284); kept in torch for callers outside the kernels.  Same order of
    operations as the reference (its constants q, A take the dtype of ``final_state_mass``, an fp32 tensor in PhaseSpace)."""
    m = -1
    q = -torch.log((final_state_mass / E_cm) ** 2)
    A = (m / 2) * q ** 2 + q ** 2
    l1 = (-q / A + torch.sqrt((-q / A) ** 2 + 2 * m * unif[:, 0] / A)) / (m / A)
    l2 = (m * l1 + q) * unif[:, 1]
    x1, x2 = torch.exp(-l1), torch.exp(-l2)
    return x1, x2, 1 / (1 / (A * x1 * x2))


def get_uniform_from_x1x2(x1, x2, final_state_mass, E_cm):
    """(x1, x2) -> uniform pair and its Jacobian (:287-309)."""
    m = -1
    q = -torch.log((final_state_mass / E_cm) ** 2)
    A = (m / 2) * q ** 2 + q ** 2
    l1, l2 = -torch.log(x1), -torch.log(x2)
    u = torch.cat(((((m / 2) * l1 ** 2 + q * l1) / A).unsqueeze(1), (l2 / (m * l1 + q)).unsqueeze(1)), axis=1)
    return u, 1 / (A * x1 * x2)
