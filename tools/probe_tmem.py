#!/usr/bin/env python
"""TMEM bandwidth probe (csrc/tc_probe.cu, tmem_bw_probe_kernel): cycles of one hidden-layer-epilogue-like pass over a
128-column fp32 accumulator per warp (32 lanes x 128 columns = 16 KB read, 16 KB written), alone and next to a stream of
TS-form MMAs."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from memflow_msci_project_b200 import _cabi
dev = torch.device("cuda:0")
a = torch.zeros(16, device=dev); w = torch.zeros(16, device=dev)
names = {1: "ld", 2: "st", 4: "split-arith", 8: "+MMAs"}
for variant in (1, 2, 3, 4, 5, 7, 9, 10, 11, 15):
    out = torch.zeros(64, device=dev)
    for rep in range(2):
        rc = _cabi.probe_lib().mf_tc_probe_gemm(_cabi._ptr(a), _cabi._ptr(w), _cabi._ptr(out), ctypes.c_int32(208), ctypes.c_int32(16),
                                                ctypes.c_int32(400 + variant), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, _cabi.probe_lib().mf_last_error()
    torch.cuda.synchronize()
    o = out.cpu().tolist()
    what = " ".join(n for b, n in names.items() if variant & b)
    print(f"variant {variant:2d} [{what:28s}]: cycles per 128-column pass, warps 0-3 = {o[0]:.0f} {o[1]:.0f} {o[2]:.0f} {o[3]:.0f}"
          + (f";  concurrent MMAs: {o[8]:.1f} cycles per MMA ({int(o[9])} issued)" if variant & 8 else ""))
