#!/usr/bin/env python
"""Cycle-level trace of the tensor-core flow kernel (needs a build with MF_TC_TRACE=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from memflow_msci_project_b200 import flows, _cabi
dev = torch.device("cuda:0")
torch.manual_seed(7)
flow = flows.NSF(features=3, context=65, transforms=5, bins=16, hidden_features=[128] * 4, passes=3,
                 base=flows.BoxUniform, base_args=[-torch.ones(3), torch.ones(3)], univariate_kwargs={"bound": 1.0},
                 gemm=sys.argv[1] if len(sys.argv) > 1 else "tc_3xf16").to(dev)
n = 148 * 256 * 4
x = (0.3 * torch.randn(n, 3, device=dev)).clamp_(-0.99, 0.99)
c = torch.randn(n, 65, device=dev)
m = flow
desc, blob = m.packed(dev, torch.float32, gemm=m.gemm)
z = torch.empty_like(x); ladj = torch.empty(n, device=dev); lp = torch.empty(n, device=dev)
bins = torch.zeros((5, n, 3), dtype=torch.int8, device=dev)
inside = torch.zeros((5, n, 3), dtype=torch.uint8, device=dev)
for _ in range(2):
    inside.zero_()
    _cabi.flow_log_prob(desc, blob, x, c, lp, z, ladj, bins, inside, m.gemm)
torch.cuda.synchronize()
tr = inside.view(-1)[: 8000 * 8].view(torch.int64).cpu().view(-1, 2)
ev = [(int(t), int(i)) for t, i in tr.tolist() if 100 <= i < 600 and t > 0]
ev.sort()
t0 = ev[0][0]
names = {1: "MMA t0 ready ", 11: "MMA t1 ready ", 2: "MMA t0 issued", 21: "MMA t1 issued", 3: "EPI t0 acc_full", 4: "EPI t0 done", 5: "EPI t0 waits"}
for t, i in ev[:90]:
    kind, ji = (i // 100, i % 10) if i < 300 else (i // 100, i % 100)
    tile = (i % 100) // 10 if i < 300 else 0
    print(f"{t - t0:8d}  id={i:4d}")
