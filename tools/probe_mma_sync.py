#!/usr/bin/env python
"""Legacy mma.sync.m16n8k8 TF32 rate on one SM (cycles per MMA per scheduler -> dense TFLOP/s per GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
from memflow_msci_project_b200 import _cabi
dev = torch.device("cuda:0")
a = torch.zeros(128, 16, device=dev); w = torch.zeros(16, 16, device=dev)
out = torch.zeros(128, 256, device=dev)
for f16 in (0, 1):
  for wps in (1, 2, 4):
      for _ in range(2):
          rc = _cabi.probe_lib().mf_tc_probe_gemm(_cabi._ptr(a), _cabi._ptr(w), _cabi._ptr(out), ctypes.c_int32(256), ctypes.c_int32(16),
                                            ctypes.c_int32(300 + 10 * f16 + wps), _cabi._stream(dev))
          _cabi.check(rc, "probe")
      torch.cuda.synchronize()
      c = out.view(-1)[0].item()   # cycles per MMA per scheduler (4 schedulers per SM)
      flop_clk_sm = 4 * 2 * 16 * 8 * (16 if f16 else 8) / c
      print(f"{'m16n8k16 fp16' if f16 else 'm16n8k8 tf32'} warps/scheduler={wps}: {c:6.2f} clk per MMA per scheduler -> {flop_clk_sm:7.0f} flop/clk/SM -> {flop_clk_sm * 148 * 1.965e9 / 1e12:6.1f} TFLOP/s at 1965 MHz")
