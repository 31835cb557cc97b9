#!/usr/bin/env python
"""Generate tests/golden/mmd.npz by running the UNMODIFIED reference MMD in this container.

memflow/unfolding_flow/mmd_loss.py (loaded by path from /root/reference; build container only) on seeded samples,
fp64 value + autograd gradients, and the fp32 value for the record.

    python tools/make_golden_mmd.py
"""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/memflow/unfolding_flow/mmd_loss.py"
CASES = [(1, 1), (2, 5), (7, 3), (64, 1), (129, 10), (200, 12), (257, 16), (50, 32), (40, 17)]


def main():
    spec = importlib.util.spec_from_file_location("ref_mmd_loss", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(20260)
    out = {"cases": np.array(CASES)}
    for B, D in CASES:
        x = torch.randn(B, D, generator=g, dtype=torch.float64)
        y = 1.2 * torch.randn(B, D, generator=g, dtype=torch.float64) + 0.3
        if B == 7:                       # logit-scaled phase-space-like values far from the origin
            x, y = 4.0 * x - 3.0, 4.0 * y - 3.0
        tag = f"B{B}_D{D}"
        out[tag + "_x"], out[tag + "_y"] = x.numpy(), y.numpy()
        for kernel in ("multiscale", "rbf"):
            xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            v = mod.MMD(xr, yr, kernel, torch.device("cpu"), torch.float64)
            v.backward()
            out[f"{tag}_{kernel}_value"] = v.detach().numpy()
            out[f"{tag}_{kernel}_gx"], out[f"{tag}_{kernel}_gy"] = xr.grad.numpy(), yr.grad.numpy()
            out[f"{tag}_{kernel}_value_f32"] = mod.MMD(x.float(), y.float(), kernel, torch.device("cpu"), torch.float32).numpy()
    path = os.path.join(ROOT, "tests", "golden", "mmd.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
