#!/usr/bin/env python
"""Short driver for `ncu --set full`: a few launches of each hot kernel at bench-like shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from memflow_msci_project_b200 import flows
from memflow_msci_project_b200.phasespace.phasespace import PhaseSpace

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "rambo"):
    for masses in ([125.25, 172.5, 172.5, 1e-5],):
        n = len(masses)
        for compute in ("f64", "f32"):
            ps = PhaseSpace(13000, [21, 21], [0] * n, final_masses=torch.tensor(masses), dev=dev, compute=compute, check_inputs=False)
            r = torch.rand(2_000_000, 3 * n - 2, device=dev).clamp_(1e-6, 1 - 1e-6)
            for _ in range(2):
                mom, w, x1, x2 = ps.get_momenta_from_ps(r)
            fin = mom[:, 2:].contiguous()
            for _ in range(2):
                ps.generator.getPSpoint_batch(fin, x1, x2, ensure_CM=False)
if which in ("all", "flow"):
    torch.manual_seed(7)
    flow = flows.NSF(features=3, context=65, transforms=5, bins=16, hidden_features=[128] * 4, passes=3,
                     base=flows.BoxUniform, base_args=[-torch.ones(3), torch.ones(3)], univariate_kwargs={"bound": 1.0}).to(dev)
    n_obj = 148 * 128 * 8
    x = (0.3 * torch.randn(n_obj, 3, device=dev)).clamp_(-0.99, 0.99)
    c = torch.randn(n_obj, 65, device=dev)
    with torch.no_grad():
        for _ in range(3):
            flow(c).log_prob(x)
torch.cuda.synchronize()
print("done")
