"""Host-side construction of ``mf_rambo_desc`` (include/memflow_b200.h).

The reference keeps the final-state masses in a torch tensor (fp32 unless the caller hands
in something else, ``memflow/phasespace/rambo_generator.py:57-58``) and therefore evaluates
every mass-only constant of the map in that dtype.  To reproduce its numbers exactly the
constants are computed here with the *same torch ops on the same dtype* and passed to the
device code as doubles:

  * ``tot_final_state_masses``                    rambo_generator.py:58
  * ``q``, ``A`` of the x1/x2 map                 phasespace/utils.py:266-269, 294-297
  * ``-q/A``, ``(-q/A)**2``, ``m/A``              phasespace/utils.py:271-272
  * ``flip(cumsum(flip(m)))``                     rambo_generator.py:484-486
  * ``torch.sum(m[j:])``                          rambo_generator.py:660
  * ``m_i**2`` (``set_square_t`` argument)        rambo_generator.py:337,339
  * ``(m_{n-1} +- m_{n-2})**2``                   rambo_generator.py:489-493, 145-149
"""
import ctypes
import math

import torch

MF_MAX_FINAL = 16
MF_QUIRK_F32_VOLUME = 1


class RamboDesc(ctypes.Structure):
    """ctypes mirror of ``struct mf_rambo_desc``."""

    _fields_ = [
        ("n_final", ctypes.c_int32),
        ("quirks", ctypes.c_int32),
        ("sqrt_s", ctypes.c_double),
        ("mass", ctypes.c_double * MF_MAX_FINAL),
        ("mass_sq", ctypes.c_double * MF_MAX_FINAL),
        ("tail_sum_fwd", ctypes.c_double * MF_MAX_FINAL),
        ("tail_sum_inv", ctypes.c_double * MF_MAX_FINAL),
        ("sum_mass", ctypes.c_double),
        ("x_q", ctypes.c_double),
        ("x_A", ctypes.c_double),
        ("x_c1", ctypes.c_double),
        ("x_c1sq", ctypes.c_double),
        ("x_c2", ctypes.c_double),
        ("last_plus_sq", ctypes.c_double),
        ("last_minus_sq", ctypes.c_double),
        ("vol_coef", ctypes.c_double),
        ("vol_fact", ctypes.c_double),
        ("pt_min_cut", ctypes.c_double),
        ("delr_min_cut", ctypes.c_double),
        ("rap_max_cut", ctypes.c_double),
    ]


def build_rambo_desc(masses, collider_energy, *, sum_mass=None, quirks=MF_QUIRK_F32_VOLUME,
                     pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1):
    """Fill a :class:`RamboDesc`.

    ``masses``: 1-D torch tensor in the order the map consumes them (already permuted for the
    inverse).  ``sum_mass``: the tensor the x1/x2 map is normalised with (defaults to
    ``torch.sum(masses)``; ``get_PS`` passes ``target_mass.sum()``).
    """
    m = torch.as_tensor(masses).detach().to("cpu").reshape(-1).clone()
    if not m.dtype.is_floating_point:
        m = m.to(torch.get_default_dtype())
    n = int(m.numel())
    if not 3 <= n <= MF_MAX_FINAL:
        # the reference itself fails for n < 3 (bisect_vec_batch returns None, rambo_generator.py:402-403)
        raise ValueError(f"n_final={n} outside 3..{MF_MAX_FINAL}")
    d = RamboDesc()
    d.n_final = n
    d.quirks = int(quirks)
    d.sqrt_s = float(collider_energy)
    tot = torch.sum(m) if sum_mass is None else torch.as_tensor(sum_mass).detach().to("cpu")
    tail_fwd = torch.flip(torch.cumsum(torch.flip(m, (-1,)), -1), (-1,))
    m2d = m.unsqueeze(0)
    for i in range(n):
        d.mass[i] = float(m[i])
        d.mass_sq[i] = float(m[i] ** 2)
        d.tail_sum_fwd[i] = float(tail_fwd[i])
        d.tail_sum_inv[i] = float(torch.sum(m2d[:, i:]))
    d.sum_mass = float(tot)
    logtau = torch.log((tot / collider_energy) ** 2)
    mm = -1
    q = -logtau
    A = (mm / 2) * q**2 + q**2
    d.x_q = float(q)
    d.x_A = float(A)
    d.x_c1 = float(-q / A)
    d.x_c1sq = float((-q / A) ** 2)
    d.x_c2 = float(mm / A)
    d.last_plus_sq = float((m[n - 1] + m[n - 2]) ** 2)
    d.last_minus_sq = float((m[n - 1] - m[n - 2]) ** 2)
    d.vol_coef = math.pow(2 * math.pi, 4 - 3 * n) * math.pow(math.pi / 2.0, n - 1)
    d.vol_fact = float(math.factorial(n - 1) * math.factorial(n - 2))
    d.pt_min_cut = float(pT_mincut)
    d.delr_min_cut = float(delR_mincut)
    d.rap_max_cut = float(rap_maxcut)
    return d
