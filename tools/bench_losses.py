#!/usr/bin/env python
"""Periodic-phi regression loss on one B200: mf_periodic_loss behind ``loss_fn_periodic`` next to the eager
torch chain the reference script runs (an in-file restatement of scripts/run_model_labframe_sampling.py:65-77 -- the
script itself cannot be imported on the GPU box), forward + backward, CUDA-event times.

    python tools/bench_losses.py
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memflow_msci_project_b200 import _cabi  # noqa: E402
from memflow_msci_project_b200.unfolding_flow.losses import loss_fn_periodic  # noqa: E402

PI = math.pi


def eager_chain(inp, target, loss_fn, device):
    inp = inp * 1.0                                  # the scripts pass an expression; it is wrapped in place
    inp[inp > PI] = inp[inp > PI] - 2 * PI
    inp[inp < -PI] = inp[inp < -PI] + 2 * PI
    d = target - inp
    d = torch.where(d > PI, d - 2 * PI, d)
    d = torch.where(d <= -PI, d + 2 * PI, d)
    return loss_fn(d, torch.zeros(d.shape, device=device, dtype=d.dtype))


def timed(fn, iters=50, warmup=10):
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
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    fn = torch.nn.HuberLoss(delta=1.0)
    for n in (4096, 1 << 20, 1 << 24):
        for dtype in (torch.float32, torch.float64):
            inp = (2.5 * torch.randn(n, generator=g, dtype=dtype)).to(dev).requires_grad_(True)
            tgt = ((2 * torch.rand(n, generator=g, dtype=dtype) - 1) * PI).to(dev)

            def run(f):
                def go():
                    inp.grad = None
                    v = f(inp, tgt, fn, dev)
                    v.backward()
                    return v
                return go

            n0 = _cabi.launch_count()
            v = run(loss_fn_periodic)()
            launches = _cabi.launch_count() - n0
            gi = inp.grad.clone()
            ve = run(eager_chain)()
            es = 4 if dtype == torch.float32 else 8
            ms = timed(run(loss_fn_periodic))
            line = {"op": "loss_fn_periodic Huber mean, value+grad", "n": n, "dtype": str(dtype).split(".")[-1],
                    "ms_b200": ms, "launches": launches, "ms_eager_torch_chain": timed(run(eager_chain)),
                    "value_abs_diff": abs(v.item() - ve.item()), "grad_max_abs_diff": (gi - inp.grad).abs().max().item(),
                    "algorithmic_GBps_fwd_bwd": n * es * 5 / (ms * 1e-3) / 1e9}
            line["speedup"] = line["ms_eager_torch_chain"] / line["ms_b200"]
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
