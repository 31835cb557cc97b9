#!/usr/bin/env python
"""One UF log_prob call on the wide mma.sync flow kernel (profiling target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from memflow_msci_project_b200 import flows
dev = torch.device("cuda:0")
torch.manual_seed(7)
flow = flows.NSF(features=10, context=16, transforms=6, bins=40, hidden_features=[512] * 2, passes=2, randperm=True,
                 base=flows.DiagNormal, base_args=[torch.zeros(10), 0.35 * torch.ones(10)], univariate_kwargs={"bound": 1.1},
                 gemm="mma_3xf16").to(dev)
n = 148 * 64 * 4
x = (0.3 * torch.randn(n, 10, device=dev)).clamp_(-0.99, 0.99)
c = torch.randn(n, 16, device=dev)
with torch.no_grad():
    for _ in range(2):
        lp = flow(c).log_prob(x)
torch.cuda.synchronize()
print("ok", float(lp.mean()))
