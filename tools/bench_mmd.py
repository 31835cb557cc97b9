#!/usr/bin/env python
"""MMD loss on one B200: mf_mmd (value + gradient, csrc/mmd.cu) next to the UNMODIFIED reference
function (baseline/_ref/memflow/unfolding_flow/mmd_loss.py, torch.mm Gram matrices + autograd) on the same GPU.

    python tools/bench_mmd.py [--batch 4096] [--dim 10]

Prints one JSON line per (kernel, dtype): CUDA-event times of forward + backward, and the value difference.
"""
import argparse
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memflow_msci_project_b200 import _cabi  # noqa: E402
from memflow_msci_project_b200.unfolding_flow.mmd_loss import MMD  # noqa: E402


def load_reference():
    path = os.path.join(ROOT, "baseline", "_ref", "memflow", "unfolding_flow", "mmd_loss.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_mmd_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MMD


def timed(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--dim", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ref = load_reference()
    g = torch.Generator().manual_seed(1)
    for dtype in (torch.float32, torch.float64):
        x = torch.randn(args.batch, args.dim, generator=g, dtype=dtype).to(dev).requires_grad_(True)
        y = (0.9 * torch.randn(args.batch, args.dim, generator=g, dtype=dtype) + 0.1).to(dev)
        for kernel in ("multiscale", "rbf"):
            def ours():
                x.grad = None
                v = MMD(x, y, kernel, dev, dtype)
                v.backward()
                return v

            def theirs():
                x.grad = None
                v = ref(x, y, kernel, dev, dtype)
                v.backward()
                return v

            n0 = _cabi.launch_count()
            v = ours()
            launches = _cabi.launch_count() - n0
            gx = x.grad.clone()
            line = {"op": "MMD value+grad", "kernel": kernel, "dtype": str(dtype).split(".")[-1], "batch": args.batch,
                    "dim": args.dim, "ms_b200": timed(ours), "launches": launches,
                    "pairs_per_s": (2 * args.batch) ** 2 / (timed(ours) * 1e-3)}
            if ref is not None:
                vr = theirs()
                line["ms_reference_torch_same_gpu"] = timed(theirs)
                line["value_abs_diff"] = abs(v.item() - vr.item())
                line["grad_max_abs_diff"] = (gx - x.grad).abs().max().item()
                line["speedup"] = line["ms_reference_torch_same_gpu"] / line["ms_b200"]
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
