#!/usr/bin/env python
"""ncu target: RAMBO forward / inverse, n = 8 (semileptonic ttH), throughput mode (fp32 arithmetic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from memflow_msci_project_b200.phasespace.phasespace import PhaseSpace
dev = torch.device("cuda:0")
masses = [4.7, 4.7, 4.7, 1e-3, 1e-3, 4.7, 1e-3, 1e-3]
ps = PhaseSpace(13000, [21, 21], [0] * 8, final_masses=torch.tensor(masses), dev=dev, compute="f32", check_inputs=False)
r = torch.rand(2_000_000, 22, device=dev).clamp_(1e-6, 1 - 1e-6)
for _ in range(2):
    mom, w, x1, x2 = ps.get_momenta_from_ps(r)
fin = mom[:, 2:].contiguous()
for _ in range(2):
    ps.generator.getPSpoint_batch(fin, x1, x2, order=list(range(8)), ensure_CM=False)
torch.cuda.synchronize()
print("done")
