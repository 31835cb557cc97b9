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
tr = inside.view(-1)[: 4000 * 16].view(torch.int64).cpu().view(-1, 2)
ev = [(int(c), int(i)) for c, i in tr.tolist() if 0 < i < 11000 and c > 0]
ev.sort()
# the kernel's own debug output (in-range flags of the first objects) lands in the same buffer: drop entries whose
# "clock" is not near the others
med = ev[len(ev) // 2][0]
ev = [e for e in ev if abs(e[0] - med) < 10 ** 9]
t0 = ev[0][0]
# id = warp * 1000 + code * 20 + job (flow_tc.cu, MF_TRACE).  One column per traced warp: the issuer (warp 1), then the
# (two issuers since r2t: warps 9 and 10), then the four epilogue warps of tile slot 0 (warps 0-3) and of slot 1 (warps 4-7).
CODES = {1: "W", 2: "rdy", 3: "iss", 4: "st", 5: "wait", 6: "full", 7: "done", 8: "tile", 9: "staged", 10: "xdone", 11: "ctx0", 12: "ctx1"}
cols = [9, 10, 0, 1, 2, 3, 4, 5, 6, 7]
print("# clk | issuer of slot 0 | issuer of slot 1 | slot 0 warps 0 1 2 3 | slot 1 warps 4 5 6 7   [W weights landed, rdy deps met, iss issued;")
print("#       wait = waiting for the accumulator, full = accumulator complete, done = epilogue finished and arrived]")
for c, i in ev[:int(sys.argv[2]) if len(sys.argv) > 2 else 900]:
    w, code, job = i // 1000, (i % 1000) // 20, i % 20
    line = ["        "] * len(cols)
    if w in cols:
        line[cols.index(w)] = f"{CODES.get(code, code)}{job:<2d}".ljust(8)
    print(f"{c - t0:8d} " + "".join(line))
