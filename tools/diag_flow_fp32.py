"""GPU diagnostic: where does the fp32 flow kernel deviate from the fp64 oracle, and is the numpy-fp32
oracle (the reference's arithmetic class) equally far?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import spec_from_module
from memflow_msci_project_b200 import flows
from oracle import flow_oracle as fo
import test_flow_gpu as T

dev = torch.device("cuda:0")
for name in ["TF-A", "TF-B", "UF"]:
    for mode in ["default", "bias", "stress"]:
        ctor, kw = T.CONFIGS[name]
        torch.manual_seed(7)
        flow = ctor(**kw)
        with torch.no_grad():
            for t in flow.core_transforms():
                lin = t.hyper.linears()[-1]
                if mode == "stress":
                    lin.weight.mul_(3.0)
                if mode in ("bias", "stress"):
                    lin.bias.uniform_(-1.5, 1.5)
        flow = flow.to(dev)
        N = 1500 if name == "UF" else 3000
        x, c = T.inputs(flow, N, 11, dev, torch.float32)
        with torch.no_grad():
            lp, z, ladj, bins, inside = flow(c).log_prob_debug(x)
        spec = spec_from_module(flow)
        xn, cn = x.double().cpu().numpy(), c.double().cpu().numpy()
        z64, l64, dbg = fo.flow_forward(spec, xn, cn, return_debug=True)
        z32, l32 = fo.flow_forward(spec, xn, cn, dtype=np.float32)
        eg = np.abs(ladj.double().cpu().numpy() - l64) / np.maximum(1, np.abs(l64))
        eo = np.abs(l32.astype(np.float64) - l64) / np.maximum(1, np.abs(l64))
        zg = np.abs(z.double().cpu().numpy() - z64).max(-1)
        zo = np.abs(z32.astype(np.float64) - z64).max(-1)
        q = lambda a: "/".join(f"{v:.1e}" for v in np.quantile(a, [0.5, 0.99, 0.999, 1.0]))
        print(f"{name:5s} {mode:8s} ladj err GPU {q(eg)} | numpy32 {q(eo)} || z err GPU {q(zg)} | numpy32 {q(zo)}")
        i = int(np.argmax(eg))
        k_ref = np.stack([d[0] for d in dbg])[:, i]; m_ref = np.stack([d[1] for d in dbg])[:, i]
        print(f"      worst obj {i}: x={xn[i]}, bins gpu {bins[:, i].cpu().numpy().tolist()} ref {k_ref.tolist()} inside ref {m_ref.astype(int).tolist()} numpy32 err there {eo[i]:.1e}")
