#!/usr/bin/env python
"""RAMBO forward + backward under autograd (generateKinematics_batch(..., requires_grad=True)): the native backward
kernel (mf_rambo_forward_backward) against the round-1 backward (torch autograd through the fp64 restatement)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from memflow_msci_project_b200.phasespace.phasespace import PhaseSpace
from memflow_msci_project_b200.phasespace import _autograd as ag

dev = torch.device("cuda:0")
for name, masses in (("n4", [125.25, 172.5, 172.5, 1e-5]), ("n8", [4.7, 4.7, 4.7, 1e-3, 1e-3, 4.7, 1e-3, 1e-3])):
    n = len(masses)
    ps = PhaseSpace(13000, [21, 21], [0] * n, final_masses=torch.tensor(masses), dev=dev)
    N = 1 << 18
    torch.manual_seed(1)
    r0 = (torch.rand(N, 3 * n - 2, device=dev, dtype=torch.float64) * 0.98 + 0.01)
    cot = torch.randn(N, n + 2, 4, device=dev, dtype=torch.float64)
    for native in (True, False):
        ag.NATIVE_BACKWARD = native
        def step():
            r = r0.clone().requires_grad_(True)
            mom, w, x1, x2 = ps.get_momenta_from_ps(r, requires_grad=True)
            ((mom * cot).sum() + x1.sum() + x2.sum()).backward()
            return r.grad
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            g = step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"rambo forward+backward {name} fp64 N={N} backward={'native kernel' if native else 'torch restatement'}: "
              f"{ms:8.3f} ms  {N / ms * 1e3:.3e} events/s", flush=True)
ag.NATIVE_BACKWARD = True
