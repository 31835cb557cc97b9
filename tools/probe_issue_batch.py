#!/usr/bin/env python
"""tcgen05.mma (TS form, one elected thread, K-block unrolled) as a function of the batch length and of the operands:
fixed cost of a short MMA sequence, zero vs random operands, one SM vs all SMs busy."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from memflow_msci_project_b200 import _cabi

dev = torch.device("cuda:0")
a = torch.zeros(128, 16, device=dev)
w = torch.zeros(16, 16, device=dev)
out = torch.zeros(128, 256, device=dev)


def run(n, count, variant, all_sms=False):
    """variant 1: zero operands, 3: random fp16 operands; mode = 100 + 20 * variant + 10 (A in TMEM) + 1 issuer."""
    mode = 100 + 20 * variant + 10 + 1 + (1000 if all_sms else 0)
    for _ in range(2):
        rc = _cabi.probe_lib().mf_tc_probe_gemm(_cabi._ptr(a), _cabi._ptr(w), _cabi._ptr(out), ctypes.c_int32(count),
                                          ctypes.c_int32(n), ctypes.c_int32(mode), _cabi._stream(dev))
        _cabi.check(rc, "mf_tc_probe_gemm")
    torch.cuda.synchronize()
    issue, done = out.view(-1)[:2].tolist()
    return issue, done


for n in (48, 96, 128):
    for count in (16, 48, 96, 192):
        issue, done = run(n, count, 1)
        print(f"zero operands  one SM   N={n:3d} count={count:3d}: issued after {issue:6.0f} clk, complete after {done:6.0f} clk "
              f"({done / count:5.1f} per MMA)")
for label, variant, all_sms in (("random operands one SM ", 3, False), ("random operands all SMs", 3, True)):
    issue, done = run(128, 192, variant, all_sms)
    print(f"{label} N=128 count=192: complete after {done:6.0f} clk ({done / 192:5.1f} per MMA)")
