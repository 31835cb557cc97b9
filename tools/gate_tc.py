#!/usr/bin/env python
"""Hang gate for the tensor-core flow kernel: a few small launches (odd tile counts, one tile, ragged tail) that a
deadlocked barrier protocol would not survive; run under a short `timeout` before anything longer (tools/gpu_iter.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from memflow_msci_project_b200 import flows
dev = torch.device("cuda:0")
torch.manual_seed(3)
flow = flows.NSF(features=3, context=65, transforms=5, bins=16, hidden_features=[128] * 4, passes=3,
                 base=flows.BoxUniform, base_args=[-torch.ones(3), torch.ones(3)], univariate_kwargs={"bound": 1.0},
                 gemm="tc_3xf16").to(dev)
ref = None
for n in (77, 128, 3 * 128 + 5, 148 * 256 + 128 * 3 + 9, 148 * 256 * 3 + 77):
    x = (0.3 * torch.randn(n, 3, device=dev)).clamp_(-0.99, 0.99)
    c = torch.randn(n, 65, device=dev)
    with torch.no_grad():
        lp = flow(c).log_prob(x)
        s = flow(c).sample()
    torch.cuda.synchronize()
    assert torch.isfinite(lp).all() and torch.isfinite(s).all(), n
    print(n, float(lp.mean()), flush=True)
print("gate ok")
