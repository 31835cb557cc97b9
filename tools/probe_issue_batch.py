import os, sys
sys.path.insert(0, "/root/repo")
import ctypes, torch
from memflow_msci_project_b200 import _cabi
dev = torch.device("cuda:0")
a = torch.zeros(128, 16, device=dev); w = torch.zeros(16, 16, device=dev)
out = torch.zeros(128, 256, device=dev)
for variant in (3, 7):
  for n in (128,):
    for count in (192,):
        for _ in range(2):
          rc = _cabi.lib().mf_tc_probe_gemm(_cabi._ptr(a), _cabi._ptr(w), _cabi._ptr(out), ctypes.c_int32(count), ctypes.c_int32(n),
                                              ctypes.c_int32(100 + 20 * (variant & 3) + 10 + 1 + (1000 if variant > 3 else 0)), _cabi._stream(dev))
          _cabi.check(rc, "probe")
        torch.cuda.synchronize()
        o = out.view(-1)[:2].tolist()
        print(f"v{variant & 3}{' all SMs' if variant > 3 else ' one SM'} TS N={n} count={count}: issue done {o[0]:.0f} clk, all complete {o[1]:.0f} clk ({o[1]/count:.1f}/MMA)")
