#!/usr/bin/env python
"""Generate tests/golden/rambo_*.npz by running the UNMODIFIED reference in this container.

Imports memflow.phasespace / memflow.unfolding_flow.utils from /root/reference (read-only,
only present in the build container) with sys.modules stubs for the packages it imports
but does not use on this path (`particle`, `vector`, `awkward`), feeds it seeded fp32
uniforms cast to fp64 (SURVEY.md section 7 parity protocol) and stores inputs + outputs.
It also cross-checks the C oracle against the reference and prints the deviations.

    python tools/make_golden_rambo.py            # writes tests/golden/*.npz
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("MEMFLOW_REFERENCE", "/root/reference")


def import_reference():
    for name in ("particle", "vector", "awkward"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["particle"].Particle = object
    sys.path.insert(0, REF)
    from memflow.phasespace.phasespace import PhaseSpace
    from memflow.unfolding_flow.utils import Compute_ParticlesTensor
    return PhaseSpace, Compute_ParticlesTensor


CASES = {
    # name: (masses, N, seed)
    "n4_HttG": ([125.25, 172.5, 172.5, 1e-5], 768, 1234),                       # scripts/run_model_labframe_sampling.py:259
    "n3_Htt": ([125.25, 172.5, 172.5], 512, 1235),                              # notebooks/TestMadgraphChain.ipynb cell 4
    "n8_ttH_decay": ([4.7, 4.7, 4.7, 1e-3, 1e-3, 4.7, 1e-3, 1e-3], 768, 1236),  # BASELINE config 1
    "n5_HttGG": ([125.25, 172.5, 172.5, 1e-5, 1e-5], 384, 1237),                # ttH + two ISR partons
    "n6_tt_decay": ([80.4, 4.7, 80.4, 4.7, 1e-5, 125.25], 384, 1238),          # W b W b g H
    # n=2 is not a reference case: bisect_vec_batch returns None for zero mass columns
    # (rambo_generator.py:402-403) and generateIntermediatesMassless_batch then raises.
}


def uniforms(N, ndim, seed, eps=1e-6):
    g = torch.Generator().manual_seed(seed)
    r = torch.rand(N, ndim, generator=g, dtype=torch.float32)
    return r.clamp_(eps, 1 - eps)


def main():
    PhaseSpace, CPT = import_reference()
    from memflow_msci_project_b200.phasespace._desc import build_rambo_desc
    from oracle import rambo as orc
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cpu")
    for name, (masses, N, seed) in CASES.items():
        n = len(masses)
        mt = torch.tensor(masses)  # fp32, as every reference call site builds it
        ps = PhaseSpace(13000, [21, 21], [0] * n, final_masses=mt, dev=dev)
        r32 = uniforms(N, 3 * n - 2, seed)
        if N == 4:
            raise SystemExit("reference breaks on a batch of exactly 4 events")
        mom, w, x1, x2 = ps.get_momenta_from_ps(r32.double())
        with contextlib.redirect_stdout(io.StringIO()):
            # PhaseSpace.get_ps_from_momenta hard-codes order=[0,1,2,3] (n=4): call the generator
            # with the identity order of the right length for the other multiplicities.
            rinv, prob = ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2, order=list(range(n)))
        out = dict(masses=np.asarray(masses, dtype=np.float32), r=r32.numpy(), momenta=mom.numpy(),
                   weight=w.numpy(), x1=x1.numpy(), x2=x2.numpy(), r_inv=rinv.numpy(),
                   detjacinv=prob.numpy())
        # cuts (only where cheap): pT and eta and deltaR cuts on the same uniforms
        if n in (3, 4):
            _, wc, _, _ = ps.get_momenta_from_ps(r32.double(), pT_mincut=30.0, delR_mincut=0.4,
                                                 rap_maxcut=2.4)
            out["weight_cuts"] = wc.numpy()
            out["cuts"] = np.asarray([30.0, 0.4, 2.4])
        # gradients of the reference's own autograd graph (requires_grad=True path, rambo_generator.py:175)
        if n in (3, 4):
            rg = r32[:128].double().clone().requires_grad_(True)
            gm, gw_, gx1, gx2 = ps.get_momenta_from_ps(rg, requires_grad=True)
            gen_ = torch.Generator().manual_seed(seed + 100)
            cm = torch.randn(gm.shape, dtype=torch.float64, generator=gen_)
            cw = torch.randn(gw_.shape, dtype=torch.float64, generator=gen_) / gw_.detach()
            c1 = torch.randn(gx1.shape, dtype=torch.float64, generator=gen_)
            c2 = torch.randn(gx2.shape, dtype=torch.float64, generator=gen_)
            ((gm * cm).sum() + (gw_ * cw).sum() + (gx1 * c1).sum() + (gx2 * c2).sum()).backward()
            np.savez_compressed(os.path.join(outdir, f"rambo_grad_{name}.npz"), masses=np.asarray(masses, dtype=np.float32),
                                r=rg.detach().numpy(), cot_momenta=cm.numpy(), cot_weight=cw.numpy(), cot_x1=c1.numpy(),
                                cot_x2=c2.numpy(), grad_r=rg.grad.numpy())
            print(f"{name}: reference gradient fixture written, max |grad| per column {rg.grad.abs().max(0).values.numpy().round(2)}")
        # oracle cross-check
        desc = build_rambo_desc(mt, 13000)
        omom, ow, olw, ox1, ox2, ofl = orc.forward(desc, r32.double().numpy())
        fin = np.isfinite(out["weight"])
        E = np.abs(out["momenta"][:, :, 0:1]) + 1e-300
        dmom = np.max(np.abs(omom - out["momenta"]) / E)
        dw = np.max(np.abs(ow[fin] / out["weight"][fin] - 1)) if fin.any() else 0.0
        same_inf = np.array_equal(np.isfinite(ow), fin)
        oinv, oprob, _ = orc.inverse(desc, out["momenta"][:, 2:], out["x1"], out["x2"])
        dr = np.abs(oinv - out["r_inv"])
        dr = np.minimum(dr, 1 - dr)  # phi columns wrap
        pf = np.isfinite(out["detjacinv"]) & (out["detjacinv"] != 0)
        dp = np.max(np.abs(oprob[pf] / out["detjacinv"][pf] - 1)) if pf.any() else 0.0
        print(f"{name}: N={N} oracle-vs-reference  max|dP|/E={dmom:.2e}  max dW/W={dw:.2e} "
              f"inf-pattern-equal={same_inf}  max|dr|={dr.max():.2e}  max dprob={dp:.2e} "
              f"finite-weight frac={fin.mean():.3f}")
        if "weight_cuts" in out:
            d2 = build_rambo_desc(mt, 13000, pT_mincut=30.0, delR_mincut=0.4, rap_maxcut=2.4)
            _, owc, *_ = orc.forward(d2, r32.double().numpy())
            agree = np.mean((owc == 0) == (out["weight_cuts"] == 0))
            print(f"    cuts: zero-pattern agreement {agree:.4f}, pass frac {np.mean(out['weight_cuts'] != 0):.3f}")
        np.savez_compressed(os.path.join(outdir, f"rambo_{name}.npz"), **out)

    # Compute_ParticlesTensor.get_PS (n = 4): lab-frame style inputs built from the forward map
    mt = torch.tensor([125.25, 172.5, 172.5, 1e-5])
    ps = PhaseSpace(13000, [21, 21], [25, 6, -6, 21], final_masses=mt, dev=dev)
    N = 768
    r32 = uniforms(N, 10, 4321)
    mom, w, x1, x2 = ps.get_momenta_from_ps(r32.double())
    boost = torch.zeros(N, 4, dtype=torch.float64)
    boost[:, 0] = (x1 + x2) * 13000 / 2
    boost[:, 3] = (x1 - x2) * 13000 / 2
    fin4 = mom[:, 2:].clone()
    # poison a few events so that every branch of mask_problems fires
    fin4[5, 0, 1] += 0.1          # not in the CM
    boost[7, 0] = 14000.0          # x1 > 1
    boost[9, 3] = boost[9, 0] * 1.5  # x2 < 0
    fin4[11, 3, 0] *= 0.5          # broken kinematics -> r outside [0,1] or NaN
    ps_out, jac, mask = CPT.get_PS(fin4, boost)
    perm = [0, 2, 1, 3]
    tm = torch.Tensor([[125.25, 172.5, 172.5, 1e-5]])
    ps_perm, _, mask_perm = CPT.get_PS(fin4, boost, order=perm, target_mass=tm)
    desc = build_rambo_desc(tm[0], 13000, sum_mass=tm.sum())
    ops, omask = orc.get_ps(desc, fin4.numpy(), boost.numpy())
    good = ~(mask.numpy() | omask)
    d = np.abs(ops - ps_out.numpy())[good]
    d = np.minimum(d, 1 - d)
    print(f"get_PS: mask equal={np.array_equal(mask.numpy(), omask)} ({mask.sum().item()} flagged) "
          f"max|dps|={np.nanmax(d):.2e}")
    desc_p = build_rambo_desc(tm[0, perm], 13000, sum_mass=tm.sum())
    ops2, omask2 = orc.get_ps(desc_p, fin4.numpy(), boost.numpy(), perm=perm)
    good = ~(mask_perm.numpy() | omask2)
    d = np.abs(ops2 - ps_perm.numpy())[good]
    d = np.minimum(d, 1 - d)
    print(f"get_PS perm: mask equal={np.array_equal(mask_perm.numpy(), omask2)} max|dps|={np.nanmax(d):.2e}")
    np.savez_compressed(os.path.join(outdir, "rambo_get_PS.npz"), momenta=fin4.numpy(),
                        boost=boost.numpy(), ps=ps_out.numpy(), mask=mask.numpy(),
                        perm=np.asarray(perm, dtype=np.int32), ps_perm=ps_perm.numpy(),
                        mask_perm=mask_perm.numpy())


if __name__ == "__main__":
    main()
