#!/usr/bin/env python
"""mf_gemm_nt_3xf16 against torch (cuBLAS fp32) on the GEMM shapes of the TF-A training step (655 360 rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from memflow_msci_project_b200 import _cabi

dev = torch.device("cuda:0")
rows = 655360


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for name, k, n in (("layer 128->128", 128, 128), ("layer 68->128", 68, 128), ("layer 128->141", 128, 141), ("dX 141->128", 141, 128)):
    x = torch.randn(rows, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.randn(n, device=dev)
    y = torch.empty(rows, n, device=dev)
    t_nat = timed(lambda: _cabi.gemm_nt(x, w, y, bias=bias, relu=True))
    t_ref = timed(lambda: torch.relu(torch.nn.functional.linear(x, w, bias)))
    gb = (rows * (k + n) * 4) / 1e9
    print(f"forward/dX {name}: native {t_nat:.3f} ms ({gb / t_nat * 1e3:.0f} GB/s), torch {t_ref:.3f} ms", flush=True)
    gz = torch.randn(rows, n, device=dev) * 1e-6
    gw = torch.zeros(n, k, device=dev)
    s = torch.exp2(torch.floor(14.0 - torch.log2(gz.abs().amax()))).reshape(1)
    gb = torch.zeros(n, device=dev)
    t_nat = timed(lambda: _cabi.gemm_tn(gz, x, gw, k_split=592, mask=y, colsum=gb, scale_a=s))
    t_ref = timed(lambda: ((gz * (y > 0)).t() @ x, (gz * (y > 0)).sum(0)))
    print(f"dW + db {name}: native {t_nat:.3f} ms, torch {t_ref:.3f} ms", flush=True)
    if k % 4 == 0 and n <= 128:
        t_nat = timed(lambda: _cabi.gemm_tn(gz, x, gw, k_split=592, colsum=gb, scale_a=s))
        print(f"dW + db {name}, gradient already masked (bulk-copied operands): native {t_nat:.3f} ms", flush=True)
