"""Autograd bridge of the RAMBO forward map (`generateKinematics_batch(..., requires_grad=True)`,
reference call sites `scripts/run_model_labframe_sampling.py:432,709`).

The FORWARD is the fused sm_100a kernel (fp64 compute).  The BACKWARD is ONE launch of `mf_rambo_forward_backward`
(csrc/rambo_bwd.cu: forward-mode dual numbers, fp64).  The differentiable fp64 torch restatement of the algorithm of
`memflow/phasespace/rambo_generator.py:168-398` below is what that kernel is tested against (and the round-1
backward, `NATIVE_BACKWARD = False`); neither path runs on the CPU or touches oracle/.

Gradient semantics follow the reference's autograd graph (checked against gradients produced by the reference
itself, tests/golden/rambo_grad_*.npz), which is NOT the full derivative of the map:
  * the massless-map root u(r_mass) comes out of `bisect_vec_batch` (`rambo_generator.py:400-446`), whose result
    is assembled from constants and `torch.where`: the mass uniforms (columns 0..n-3 of r) get zero gradient;
  * the intermediate-mass tensor M is a fresh leaf whose first column is filled with E_cm under `no_grad`
    (`rambo_generator.py:264-284`): the outgoing momenta and the mass-dependent weight factors do not see the
    x1/x2 uniforms; only the incoming partons, x1, x2, the x1x2 Jacobian and the flat volume (E_cm^2)^(n-2) do.
`reference_quirks=False` on the generator switches to the full derivative (implicit-function derivative of the
root, du/dv = 1 / (e (e+1) u^(e-1) (1-u)), and E_cm flowing into M).
"""
import math

import torch

from .. import _cabi


# True: the backward is the native kernel (csrc/rambo_bwd.cu).  False: the differentiable torch restatement below --
# kept as the checker the kernel is tested against (tests/test_rambo_gpu.py) and as documentation of the reference's
# autograd graph.
NATIVE_BACKWARD = True


def _massless_map(u, e):
    return u**e * ((e + 1.0) - e * u)


@torch.no_grad()
def _solve_u(v, e):
    """Root u in [0,1] of (e+1) u^e - e u^(e+1) = v, per column exponent e (64 halvings, fp64)."""
    lo = torch.zeros_like(v)
    hi = torch.ones_like(v)
    for _ in range(64):
        mid = 0.5 * (lo + hi)
        below = _massless_map(mid, e) < v
        lo = torch.where(below, mid, lo)
        hi = torch.where(below, hi, mid)
    return 0.5 * (lo + hi)


def _rho(M, N, m):
    Msq = M * M
    return torch.sqrt((Msq - (N + m) ** 2) * (Msq - (N - m) ** 2)) / (8.0 * Msq)


def _set_square(p3, square):
    return torch.sqrt((p3 * p3).sum(-1) + square)


def _boost(E, p3, bv):
    """phasespace/utils.py:88-118: boost (E, p3) by the velocity bv."""
    b2 = (bv * bv).sum(-1)
    gamma = 1.0 / torch.sqrt(1.0 - b2)
    bp = (p3 * bv).sum(-1)
    gamma2 = torch.where(b2 > 0, (gamma - 1.0) / torch.where(b2 > 0, b2, torch.ones_like(b2)), torch.zeros_like(b2))
    factor = gamma2 * bp + gamma * E
    return gamma * (E + bp), p3 + factor[:, None] * bv


def eager_forward(gen, d, r, full_derivative):
    """Differentiable fp64 restatement: r [N, 3n-2] -> (momenta [N, n+2, 4], weight [N], x1 [N], x2 [N]).
    `d` is the generator's RamboDesc (mass-only constants, evaluated like the reference does)."""
    n = gen.n_final
    ua, ub = r[:, 3 * n - 4], r[:, 3 * n - 3]
    # get_x1x2_from_uniform, phasespace/utils.py:260-284
    mlogx1 = (d.x_c1 + torch.sqrt(d.x_c1sq + (-2.0 * ua) / d.x_A)) / d.x_c2
    mlogx2 = (d.x_q - mlogx1) * ub
    x1, x2 = torch.exp(-mlogx1), torch.exp(-mlogx2)
    wgt_x = d.x_A * x1 * x2
    E_cm = torch.sqrt(x1 * x2) * d.sqrt_s
    K0 = E_cm - d.sum_mass
    if not full_derivative:
        # reference graph: the intermediate-mass tensor M is a fresh leaf whose first column is filled with E_cm
        # under no_grad (rambo_generator.py:274-284), so nothing downstream of M sees the x1/x2 uniforms
        K0 = K0.detach()
    # intermediate masses, :448-509
    masses = [d.mass[i] for i in range(n)]
    if n > 2:
        e = torch.arange(n - 2, 0, -1, dtype=r.dtype, device=r.device)[None, :]
        v = r[:, : n - 2]
        u = _solve_u(v.detach(), e)
        if full_derivative:
            dv = e * (e + 1.0) * u ** (e - 1.0) * (1.0 - u)
            u = u + (v - v.detach()) / dv
        prod = torch.cumprod(torch.cat((torch.ones_like(K0)[:, None], u), dim=1), dim=1)
    else:
        prod = torch.ones_like(K0)[:, None]
    K = K0[:, None] * prod                                           # [N, n-1]
    tail = torch.tensor([d.tail_sum_fwd[i] for i in range(n - 1)], dtype=r.dtype, device=r.device)
    M = K + tail
    # weight, :111-140 and :489-507 (the fp32 roundings of the reference do not change the derivative)
    logw = math.log(d.vol_coef) + (n - 2) * torch.log(E_cm * E_cm) - math.log(d.vol_fact)
    Ml = M[:, n - 2]
    logw = logw + torch.log(8.0 * torch.sqrt((Ml * Ml - d.last_plus_sq) * (Ml * Ml - d.last_minus_sq)) / (8.0 * Ml * Ml))
    for i in range(n - 2):
        term = (_rho(M[:, i], M[:, i + 1], masses[i]) / _rho(K[:, i], K[:, i + 1], 0.0)) * (M[:, i + 1] / K[:, i + 1])
        logw = logw + torch.log(term)
    logw = logw + (2 * n - 4) * torch.log(K[:, 0] / M[:, 0]) + torch.log(wgt_x)
    weight = torch.exp(logw)
    # decay chain, :294-345
    Mfull = torch.cat((M, torch.full_like(K0, masses[n - 1])[:, None]), dim=1)
    QE = Mfull[:, 0]
    Q3 = torch.zeros((r.shape[0], 3), dtype=r.dtype, device=r.device)
    rnd = r[:, n - 2:]
    out = []
    for i in range(n - 1):
        q = 4.0 * Mfull[:, i] * _rho(Mfull[:, i], Mfull[:, i + 1], masses[i])
        cos_t = 2.0 * rnd[:, 2 * i] - 1.0
        sin_t = torch.sqrt(1.0 - cos_t * cos_t)
        phi = 2.0 * math.pi * rnd[:, 2 * i + 1]
        cos_p = torch.cos(phi)
        s = torch.sqrt(1.0 - cos_p * cos_p)
        sin_p = torch.where(phi > math.pi, -s, s)
        p3 = torch.stack((q * sin_t * cos_p, q * sin_t * sin_p, q * cos_t), dim=1)
        E = _set_square(p3, d.mass_sq[i])
        E, p3 = _boost(E, p3, Q3 / QE[:, None])
        E = _set_square(p3, d.mass_sq[i])
        out.append(torch.cat((E[:, None], p3), dim=1))
        Q3 = Q3 - p3
        QE = _set_square(Q3, Mfull[:, i + 1] ** 2)
    out.append(torch.cat((QE[:, None], Q3), dim=1))
    # incoming partons, :533-551
    z = torch.zeros_like(E_cm)
    half = E_cm / 2.0
    mom = torch.stack([torch.stack((half, z, z, half), 1), torch.stack((half, z, z, -half), 1)] + out, dim=1)
    return mom, weight, x1, x2


class RamboForwardFn(torch.autograd.Function):
    """(generator, r, cuts...) -> (momenta, weight, x1, x2): kernel forward, eager-torch recompute backward."""

    @staticmethod
    def forward(ctx, gen, r, pT_mincut, delR_mincut, rap_maxcut):
        in_dtype = r.dtype
        rd = r.detach().double().contiguous()   # the reference runs this path in fp64
        N, n = rd.shape[0], gen.n_final
        desc = gen._desc("fwd", gen.masses_t, (pT_mincut, delR_mincut, rap_maxcut))
        mom = torch.empty((N, n + 2, 4), dtype=torch.float64, device=rd.device)
        weight = torch.empty(N, dtype=torch.float64, device=rd.device)
        x1 = torch.empty(N, dtype=torch.float64, device=rd.device)
        x2 = torch.empty(N, dtype=torch.float64, device=rd.device)
        _cabi.rambo_forward(desc, rd, mom, weight, None, x1, x2, None, _cabi.MF_COMPUTE_F64)
        ctx.gen, ctx.desc, ctx.in_dtype = gen, desc, in_dtype
        ctx.save_for_backward(rd, weight)
        if gen.reference_dtypes:
            return mom, weight, x1.to(in_dtype), x2.to(in_dtype)  # x1/x2 follow the uniforms' dtype
        return mom.to(in_dtype), weight.to(in_dtype), x1.to(in_dtype), x2.to(in_dtype)

    @staticmethod
    def backward(ctx, g_mom, g_w, g_x1, g_x2):
        rd, w_fwd = ctx.saved_tensors
        if NATIVE_BACKWARD:
            # one launch of mf_rambo_forward_backward (forward-mode duals, fp64): no tape, no eager kernels
            def c64(g):
                return None if g is None else g.detach().double().contiguous()
            gr = torch.empty_like(rd)
            _cabi.rambo_forward_backward(ctx.desc, rd, c64(g_mom), c64(g_w), c64(g_x1), c64(g_x2), w_fwd,
                                         not ctx.gen.reference_quirks, gr)
            return None, gr.to(ctx.in_dtype), None, None, None
        with torch.enable_grad():
            rr = rd.detach().requires_grad_(True)
            mom, weight, x1, x2 = eager_forward(ctx.gen, ctx.desc, rr, full_derivative=not ctx.gen.reference_quirks)
            weight = weight * (w_fwd != 0).to(weight.dtype)   # cuts zero the weight (piecewise constant factor)
            outs, cots = [], []
            for o, g in ((mom, g_mom), (weight, g_w), (x1, g_x1), (x2, g_x2)):
                if g is not None:
                    outs.append(o)
                    cots.append(g.to(o.dtype))
            (gr,) = torch.autograd.grad(outs, [rr], cots)
        return None, gr.to(ctx.in_dtype), None, None, None
