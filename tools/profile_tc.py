#!/usr/bin/env python
"""ncu target: a few launches of the tcgen05 flow kernel at the bench shape (TF-A)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from memflow_msci_project_b200 import flows
dev = torch.device("cuda:0")
torch.manual_seed(7)
gemm = sys.argv[1] if len(sys.argv) > 1 else "tc_3xf16"
flow = flows.NSF(features=3, context=65, transforms=5, bins=16, hidden_features=[128] * 4, passes=3,
                 base=flows.BoxUniform, base_args=[-torch.ones(3), torch.ones(3)], univariate_kwargs={"bound": 1.0}, gemm=gemm).to(dev)
n_obj = 148 * 256 * 4
x = (0.3 * torch.randn(n_obj, 3, device=dev)).clamp_(-0.99, 0.99)
c = torch.randn(n_obj, 65, device=dev)
with torch.no_grad():
    for _ in range(3):
        flow(c).log_prob(x)
torch.cuda.synchronize()
print("done")
