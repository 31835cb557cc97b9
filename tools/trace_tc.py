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
ev = [(int(t), int(i)) for t, i in tr.tolist() if 100 <= i < 1000 and t > 0]
ev.sort()
t0 = ev[0][0]
# ids: 1TJ issuer T sees operands for job J, 2TJ issuer T issued job J; slot 0 epilogue 5JJ waits / 3JJ sees the
# accumulator / 4JJ done; slot 1 epilogue 8JJ / 6JJ / 7JJ.  Printed as two columns, one per tile slot.
def slot(i):
    if i < 300: return (i % 100) // 10
    if i >= 950: return (i - 950) // 10
    return 0 if (i < 600 or i >= 900) else 1
def label(i):
    if i < 300: return ("MMA ready  " if i < 200 else "MMA issued ") + str(i % 10)
    if i >= 950: return "   tile " + ["begin", "x staged", "ctx staged", "prefetch+bias done"][(i - 950) % 10]
    if i >= 900: return "   spline " + ["start", "ld w/h done", "scan done", "deriv sel done", "finish done"][i - 900]
    k = (i - 300 * slot(i)) // 100
    return {3: "EPI accfull ", 4: "EPI done    ", 5: "EPI wait    "}[k] + str(i % 100)
for t, i in ev[:1500]:
    print(f"{t - t0:8d}  " + ("" if slot(i) == 0 else " " * 28) + label(i))
