#!/usr/bin/env python
"""Generate tests/golden/losses.npz by executing the UNMODIFIED reference ``loss_fn_periodic``.

The function lives in a training script with heavy imports (scripts/run_model_labframe_sampling.py:65-77), so its
source is cut out of the script with ``ast`` and executed as is, with PI = torch.pi as at :46 (build container only).

    python tools/make_golden_losses.py
"""
import ast
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/scripts/run_model_labframe_sampling.py"


def reference_function(name):
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "PI": torch.pi}
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns[name]


def main():
    ref = reference_function("loss_fn_periodic")
    g = torch.Generator().manual_seed(77)
    out = {}
    n = 1537
    # predictions overshoot the [-pi, pi] range (unbounded regression output), targets are angles
    inp = 2.5 * torch.randn(n, generator=g, dtype=torch.float64)
    target = (2 * torch.rand(n, generator=g, dtype=torch.float64) - 1) * torch.pi
    inp[:8] = torch.tensor([torch.pi, -torch.pi, 3 * torch.pi, -3 * torch.pi, 0.0, 3.5, -3.5, 6.4], dtype=torch.float64)
    target[:8] = torch.tensor([-torch.pi, torch.pi, 0.0, 0.0, torch.pi, -3.0, 3.0, -0.1], dtype=torch.float64)
    out["inp"], out["target"] = inp.numpy(), target.numpy()
    dev = torch.device("cpu")
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        # the reference's zeros(...) target has torch's DEFAULT dtype; with fp64 inputs torch >= 2.2 refuses the mixed
        # Huber backward, so the fp64 run sets the default dtype the way the fp64 training configs do
        torch.set_default_dtype(dt)
        for name, fn in (("huber1", torch.nn.HuberLoss(delta=1.0)), ("huber03", torch.nn.HuberLoss(delta=0.3)),
                         ("mse", torch.nn.MSELoss()), ("huber1_none", torch.nn.HuberLoss(delta=1.0, reduction="none")),
                         ("huber1_sum", torch.nn.HuberLoss(delta=1.0, reduction="sum"))):
            a = inp.to(dt).clone().requires_grad_(True)
            t = target.to(dt).clone().requires_grad_(True)
            v = ref(a * 1.0, t, fn, dev)           # the scripts pass an expression; the function wraps it in place
            w = torch.linspace(0.5, 1.5, n, dtype=dt) if v.dim() else torch.tensor(1.0, dtype=dt)
            (v * w).sum().backward()
            out[f"{name}_{tag}_value"] = v.detach().numpy()
            out[f"{name}_{tag}_ginp"], out[f"{name}_{tag}_gtarget"] = a.grad.numpy(), t.grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "losses.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
