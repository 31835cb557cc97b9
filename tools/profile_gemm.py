#!/usr/bin/env python
"""ncu target: the forward (NT) and weight-gradient (TN) GEMMs of the training step at the TF-A hidden-layer shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from memflow_msci_project_b200 import _cabi
dev = torch.device("cuda:0")
rows, k, n = 655360, 128, 128
x = torch.randn(rows, k, device=dev)
w = torch.randn(n, k, device=dev) / k ** 0.5
b = torch.randn(n, device=dev)
y = torch.empty(rows, n, device=dev)
gz = torch.randn(rows, n, device=dev) * 1e-6
gw = torch.zeros(n, k, device=dev)
gb = torch.zeros(n, device=dev)
s = torch.exp2(torch.floor(14.0 - torch.log2(gz.abs().amax()))).reshape(1)
for _ in range(2):
    _cabi.gemm_nt(x, w, y, bias=b, relu=True)
    _cabi.gemm_tn(gz, x, gw, k_split=592, mask=y, colsum=gb, scale_a=s)
torch.cuda.synchronize()
print("ok")
