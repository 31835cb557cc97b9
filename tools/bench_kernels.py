#!/usr/bin/env python
"""Per-kernel timings on one B200 (CUDA events, warm-up, inputs larger than L2) with the algorithmic
bytes / FLOPs of SURVEY.md section 8(d) -> achieved GB/s / TFLOP/s and roofline fractions.

    python tools/bench_kernels.py [--out gpurun_out/kernels.json] [--quick]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from memflow_msci_project_b200 import flows
from memflow_msci_project_b200.phasespace.phasespace import PhaseSpace

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = PEAKS.get("hbm_gbs", 6650.0)
TF = PEAKS.get("bf16_tflops_sustained", 1400.0)


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="all", choices=["all", "rambo", "flow", "rqs"])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    res = []
    # ---------------- RAMBO
    cases = {"n4": [125.25, 172.5, 172.5, 1e-5], "n8": [4.7, 4.7, 4.7, 1e-3, 1e-3, 4.7, 1e-3, 1e-3]}
    N = 2_000_000 if args.quick else 8_000_000
    for name, masses in (cases.items() if args.only in ("all", "rambo") else []):
        n = len(masses)
        for compute in ["f64", "f32"]:
            ps = PhaseSpace(13000, [21, 21], [0] * n, final_masses=torch.tensor(masses), dev=dev, compute=compute,
                            check_inputs=False)
            r = torch.rand(N, 3 * n - 2, device=dev).clamp_(1e-6, 1 - 1e-6)
            ms = timeit(lambda: ps.get_momenta_from_ps(r))
            bytes_ev = 4 * (3 * n - 2) + 16 * (n + 2) + 12
            gbs = N * bytes_ev / (ms * 1e-3) / 1e9
            res.append(dict(kernel=f"rambo_forward {name} fp32-io compute={compute}", events=N, ms=ms, events_per_s=N / (ms * 1e-3),
                            bytes_per_event=bytes_ev, achieved_gbs=gbs, hbm_frac=gbs / HBM))
            mom, w, x1, x2 = ps.get_momenta_from_ps(r)
            fin = mom[:, 2:].contiguous()
            ms = timeit(lambda: ps.generator.getPSpoint_batch(fin, x1, x2, order=list(range(n)), ensure_CM=False))
            bytes_ev = 16 * n + 8 + 4 * (3 * n - 2) + 4
            gbs = N * bytes_ev / (ms * 1e-3) / 1e9
            res.append(dict(kernel=f"rambo_inverse {name} fp32-io compute={compute}", events=N, ms=ms, events_per_s=N / (ms * 1e-3),
                            bytes_per_event=bytes_ev, achieved_gbs=gbs, hbm_frac=gbs / HBM))
            del r, mom, fin
    # ---------------- flows (exact fp32 path)
    fl = {
        "TF-A": (dict(features=3, context=65, transforms=5, bins=16, hidden_features=[128] * 4, passes=3,
                      base=flows.BoxUniform, base_args=[-torch.ones(3), torch.ones(3)], univariate_kwargs={"bound": 1.0}), 4_000_000),
        "TF-B": (dict(features=3, context=65, transforms=4, bins=30, hidden_features=[128] * 2, passes=2, randperm=True,
                      base=flows.DiagNormal, base_args=[torch.zeros(3), 0.35 * torch.ones(3)], univariate_kwargs={"bound": 1.0}), 4_000_000),
        "UF": (dict(features=10, context=16, transforms=6, bins=40, hidden_features=[512] * 2, passes=2, randperm=True,
                    base=flows.DiagNormal, base_args=[torch.zeros(10), 0.35 * torch.ones(10)], univariate_kwargs={"bound": 1.1}), 500_000),
        # the most frequent unfolding-flow shape of the reference's configs
        "UF-common": (dict(features=10, context=32, transforms=10, bins=35, hidden_features=[128] * 4, passes=2, randperm=True,
                           base=flows.DiagNormal, base_args=[torch.zeros(10), 0.3 * torch.ones(10)], univariate_kwargs={"bound": 1.5}), 1_000_000),
        "TF-1D": (dict(features=1, context=66, transforms=5, bins=16, hidden_features=[256] * 3, passes=1,
                       base=flows.BoxUniform, base_args=[-torch.ones(1), torch.ones(1)], univariate_kwargs={"bound": 1.0}), 2_000_000),
    }
    for name, (kw, n_obj) in (fl.items() if args.only in ("all", "flow") else []):
        if args.quick:
            n_obj //= 4
        torch.manual_seed(7)
        flow = flows.NSF(**kw).to(dev)
        D, C, T, K = kw["features"], kw["context"], kw["transforms"], kw["bins"]
        H = kw["hidden_features"]
        fl_obj = T * 2 * ((D + C) * H[0] + sum(H[i] * H[i + 1] for i in range(len(H) - 1)) + H[-1] * D * (3 * K - 1))
        x = (0.3 * torch.randn(n_obj, D, device=dev)).clamp_(-0.99, 0.99)
        c = torch.randn(n_obj, C, device=dev)
        with torch.no_grad():
            for gname, gmode in (("exact", flows.modules.GEMM_EXACT), ("tc_3xf16", flows.modules.GEMM_TC_3XF16),
                                 ("tc_bf16", flows.modules.GEMM_TC_BF16), ("mma_3xf16", flows.modules.GEMM_MMA_3XF16)):
                if gmode in (flows.modules.GEMM_TC_3XF16, flows.modules.GEMM_TC_BF16) and max(H) > 128:
                    continue
                if gmode == flows.modules.GEMM_MMA_3XF16 and name not in ("UF", "TF-1D", "TF-A", "UF-common"):
                    continue
                flow.gemm = gmode
                ms = timeit(lambda: flow(c).log_prob(x), reps=5)
                tf = fl_obj * n_obj / (ms * 1e-3) / 1e12
                res.append(dict(kernel=f"flow log_prob {name} fp32 {gname}", objects=n_obj, ms=ms, objects_per_s=n_obj / (ms * 1e-3),
                                flops_per_object=fl_obj, achieved_tflops=tf, tensor_frac=tf / TF,
                                bytes_per_object=4 * (D + C) + 4, achieved_gbs=n_obj * (4 * (D + C) + 4) / (ms * 1e-3) / 1e9))
            flow.gemm = flows.modules.GEMM_EXACT
            nz = torch.randn(n_obj, D, device=dev)
            for gname, gmode in (("exact", flows.modules.GEMM_EXACT), ("tc_3xf16", flows.modules.GEMM_TC_3XF16),
                                 ("mma_3xf16", flows.modules.GEMM_MMA_3XF16)):
                if gmode == flows.modules.GEMM_TC_3XF16 and max(H) > 128:
                    continue
                if gmode == flows.modules.GEMM_MMA_3XF16 and name not in ("UF", "TF-1D", "TF-A", "UF-common"):
                    continue
                flow.gemm = gmode
                ms = timeit(lambda: flow(c).sample_from_noise(nz), reps=3)
                tf = fl_obj * kw["passes"] * n_obj / (ms * 1e-3) / 1e12
                res.append(dict(kernel=f"flow sample {name} fp32 {gname} (passes={kw['passes']})", objects=n_obj, ms=ms,
                                objects_per_s=n_obj / (ms * 1e-3), achieved_tflops=tf, tensor_frac=tf / TF))
            flow.gemm = flows.modules.GEMM_EXACT
            # fp64 flows (103 of the reference's 106 training configs run in float64): flow_kernel<double> on the fp64 pipe
            if name in ("TF-A", "UF-common"):
                n64 = n_obj // 8
                f64 = flows.NSF(**kw).to(dev).double()
                x64, c64 = x[:n64].double(), c[:n64].double()
                ms = timeit(lambda: f64(c64).log_prob(x64), reps=3)
                tf = fl_obj * n64 / (ms * 1e-3) / 1e12
                res.append(dict(kernel=f"flow log_prob {name} fp64 exact (flow_kernel<double>)", objects=n64, ms=ms,
                                objects_per_s=n64 / (ms * 1e-3), flops_per_object=fl_obj, achieved_tflops=tf,
                                tensor_frac=tf / 40.0, note="fraction of the nominal 40 TFLOP/s fp64 (no measured fp64 peak)"))
                del f64, x64, c64
        del x, c, flow
    # ---------------- standalone spline on materialised parameters (HBM-bound)
    if args.only in ("all", "rqs", "flow"):
        for K in (16, 30):
            n_el = 4_000_000 if args.quick else 16_000_000
            P = 3 * K - 1
            phi = torch.randn(n_el, P, device=dev)
            xx = (0.5 * torch.randn(n_el, device=dev)).clamp_(-0.99, 0.99)
            t = flows.MonotonicRQSTransform.from_packed(phi, K, bound=1.0)
            ms = timeit(lambda: t.call_and_ladj(xx), reps=5)
            bytes_el = 4 * P + 4 + 8
            gbs = n_el * bytes_el / (ms * 1e-3) / 1e9
            res.append(dict(kernel=f"rqs_transform standalone K={K} fp32 (fwd+ladj)", elements=n_el, ms=ms, events_per_s=n_el / (ms * 1e-3),
                            bytes_per_event=bytes_el, achieved_gbs=gbs, hbm_frac=gbs / HBM))
            del phi, xx, t
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(peaks=dict(hbm_gbs=HBM, bf16_tflops_sustained=TF), results=res), open(args.out, "w"), indent=1)
    for r in res:
        extra = f"{r.get('achieved_gbs', 0):8.1f} GB/s ({100 * r.get('hbm_frac', 0):5.1f}% HBM)" if "hbm_frac" in r else \
            f"{r['achieved_tflops']:7.2f} TFLOP/s ({100 * r['tensor_frac']:5.2f}% of bf16 sustained)"
        rate = r.get("events_per_s", r.get("objects_per_s"))
        print(f"{r['kernel']:55s} {r['ms']:9.3f} ms  {rate:10.3e} /s  {extra}")


if __name__ == "__main__":
    main()
