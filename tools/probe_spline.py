#!/usr/bin/env python
"""Register-spline throughput on one SM: cycles per K=16 spline per warp at 1..8 warps per scheduler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
from memflow_msci_project_b200 import _cabi
dev = torch.device("cuda:0")
a = torch.zeros(128, 16, device=dev); w = torch.zeros(16, 16, device=dev)
out = torch.zeros(128, 256, device=dev)
for variant in (0, 1):
    for wps in (1, 2, 4):
        for _ in range(2):
            rc = _cabi.probe_lib().mf_tc_probe_gemm(_cabi._ptr(a), _cabi._ptr(w), _cabi._ptr(out), ctypes.c_int32(256), ctypes.c_int32(16),
                                              ctypes.c_int32(200 + 10 * variant + wps), _cabi._stream(dev))
            _cabi.check(rc, "probe")
        torch.cuda.synchronize()
        c = out.view(-1)[0].item()
        print(f"variant {variant} ({'scan' if variant == 0 else 'scan+finish'}) warps/scheduler={wps}: {c:7.0f} clk per spline per warp -> {c / wps:7.0f} clk per warp-spline per scheduler")
