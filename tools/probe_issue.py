#!/usr/bin/env python
"""tcgen05.mma issue-rate probe: cycles per MMA (M=128, K=16) for one or two issuing warps, SS and TS forms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
from memflow_msci_project_b200 import _cabi
dev = torch.device("cuda:0")
a = torch.zeros(128, 16, device=dev); w = torch.zeros(16, 16, device=dev)
out = torch.zeros(128, 256, device=dev)
for variant in (0, 1):
  for n in (64, 128):
    for ts in (0, 1):
        for issuers in (1, 2):
            for count in (192,):
                out.zero_()
                for _ in range(2):
                    rc = _cabi.probe_lib().mf_tc_probe_gemm(_cabi._ptr(a), _cabi._ptr(w), _cabi._ptr(out), ctypes.c_int32(count), ctypes.c_int32(n),
                                                      ctypes.c_int32(100 + 20 * variant + 10 * ts + issuers), _cabi._stream(dev))
                    _cabi.check(rc, "probe")
                torch.cuda.synchronize()
                o = out.view(-1)[:4].tolist()
                print(f"v{variant} N={n:3d} {'TS' if ts else 'SS'} issuers={issuers} count={count}: issue {o[0]/count:6.1f} clk/MMA, done {o[1]/count:6.1f} clk/MMA"
                      + (f" | warp1 issue {o[2]/count:6.1f} done {o[3]/count:6.1f}" if issuers == 2 else ""))
