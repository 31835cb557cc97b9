#!/usr/bin/env python
"""Generate tests/golden/ps_utils.npz by running the UNMODIFIED reference helpers in this container.

memflow/phasespace/utils.py (set_square_t, rho2_t, rho2_tt, boostVector_t, square_t, mass_t, dot_t, dot_s, boost_t,
boost_tt, boost_to_lab_frame, get_x1x2_from_uniform, get_uniform_from_x1x2) and the kinematic glue of
memflow/unfolding_flow/utils.py (Compute_ParticlesTensor.get_ptetaphi_comp, get_cartesian_comp) plus the lab -> CM
-> get_PS chain of memflow/unfolding_flow/unfolding_flow.py:137-163, on seeded inputs.  The reference is imported from
/root/reference (build container only; stubs for its unused `particle`, `vector`, `awkward` imports).

    python tools/make_golden_ps_utils.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_rambo import import_reference  # noqa: E402


def main():
    PhaseSpace, CPT = import_reference()
    from memflow.phasespace import utils as ru
    g = torch.Generator().manual_seed(4242)
    N, P = 257, 4
    out = {}
    # physical four-vectors: massive particles with random momenta
    m = torch.tensor([125.25, 172.5, 172.5, 1e-5], dtype=torch.float64)
    pv = 300.0 * torch.randn(N, P, 3, generator=g, dtype=torch.float64)
    E = torch.sqrt((pv ** 2).sum(-1) + m ** 2)
    p4 = torch.cat((E.unsqueeze(-1), pv), -1)                       # [N,P,4]
    q4 = p4.sum(1)                                                   # [N,4] total
    out["p4"], out["q4"] = p4.numpy(), q4.numpy()
    out["rho2_t"] = ru.rho2_t(q4).numpy()
    out["rho2_tt"] = ru.rho2_tt(p4).numpy()
    out["boostVector_t"] = ru.boostVector_t(q4).numpy()
    out["square_t"] = ru.square_t(p4[:, 1]).numpy()
    out["mass_t"] = ru.mass_t(p4[:, 1]).numpy()
    out["dot_t"] = ru.dot_t(p4[:, 0], p4[:, 2]).numpy()
    out["dot_s"] = ru.dot_s(pv[:, 0], pv[:, 1]).numpy()
    out["set_square_t"] = ru.set_square_t(p4[:, 0], 91.19 ** 2).numpy()
    beta = ru.boostVector_t(q4)
    out["boost_t"] = ru.boost_t(p4[:, 0], -beta).numpy()
    out["boost_tt"] = ru.boost_tt(p4, -beta.unsqueeze(1)).numpy()
    x1 = torch.rand(N, generator=g, dtype=torch.float64) * 0.5 + 0.01
    x2 = torch.rand(N, generator=g, dtype=torch.float64) * 0.5 + 0.01
    inc = torch.zeros(N, 2, 4, dtype=torch.float64)
    inc[:, 0, 0] = 6500.0; inc[:, 0, 3] = 6500.0; inc[:, 1, 0] = 6500.0; inc[:, 1, 3] = -6500.0
    mom6 = torch.cat((inc, p4), 1)
    out["x1"], out["x2"], out["mom6"] = x1.numpy(), x2.numpy(), mom6.numpy()
    out["boost_to_lab_frame"] = ru.boost_to_lab_frame(mom6.clone(), x1, x2).numpy()
    unif = torch.rand(N, 2, generator=g, dtype=torch.float64)
    a, b, w = ru.get_x1x2_from_uniform(unif, torch.tensor(470.25001), 13000.0)
    out["unif"] = unif.numpy()
    out["x1x2_from_uniform"] = torch.stack((a, b, w), 1).numpy()
    u2, jac = ru.get_uniform_from_x1x2(a, b, torch.tensor(470.25001), 13000.0)
    out["uniform_from_x1x2"] = torch.cat((u2, jac.unsqueeze(1)), 1).numpy()
    # glue of unfolding_flow/utils.py
    out["ptetaphi"] = CPT.get_ptetaphi_comp(p4[:, 1]).numpy()
    pte = CPT.get_ptetaphi_comp(p4[:, 1])[:, 1:]
    out["cartesian"] = CPT.get_cartesian_comp(pte, 172.5).numpy()
    # lab -> CM -> get_PS chain (unfolding_flow.py:137-163) on a lab-frame event sample built from RAMBO points
    ps = PhaseSpace(13000, [21, 21], [25, 6, -6, 21], dev="cpu")
    r = torch.rand(N, 10, generator=g, dtype=torch.float32).clamp_(1e-4, 1 - 1e-4).double()
    mom_cm, _, xa, xb = ps.get_momenta_from_ps(r)
    lab = ru.boost_to_lab_frame(mom_cm.clone(), xa.double(), xb.double())[:, 2:]           # final state in the lab frame
    boost = lab.sum(1)
    boost[:5, 0] = 0.9 * torch.abs(boost[:5, 3])        # unphysical regressed boosts: exercised by the fix-up (:140-143)
    M_MIN_TOT = CPT.M_MIN_TOT
    boost_fix = boost.clone()
    wrong = torch.sqrt(boost_fix[:, 0] ** 2 - boost_fix[:, 3] ** 2) < M_MIN_TOT
    boost_fix[wrong, 0] = torch.sqrt(boost_fix[wrong, 3] ** 2 + M_MIN_TOT ** 2 + 1e-3)
    bv = ru.boostVector_t(boost_fix).unsqueeze(1)
    cm = ru.boost_tt(lab, -bv)
    psr, jac0, mask = CPT.get_PS(cm, boost_fix)
    out["lab"], out["lab_boost"], out["lab_boost_fixed"] = lab.numpy(), boost.numpy(), boost_fix.numpy()
    out["lab_cm"], out["lab_ps"], out["lab_mask"] = cm.numpy(), psr.numpy(), mask.numpy()
    out["m_min_tot"] = np.float64(M_MIN_TOT)
    path = os.path.join(ROOT, "tests", "golden", "ps_utils.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
