#!/usr/bin/env python
"""bench.py -- headline benchmark: events/s of flow log_prob + phase-space Jacobian (ttH, 1M events).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--events E] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one pass of the hot path over one batch of `events` synthetic ttH-shaped events PER GPU:
  1. RAMBO-on-diet forward map + Jacobian weight, n = 4 (H, t, tbar, g), uniforms -> momenta, weight, log w
  2. transfer-flow log_prob (TF-A: D=3, C=65, T=5, K=16, hidden [128]x4 -- the jets flow of
     TransferFlow_Paper_AllPartons_btag) on P = 10 objects per event, context = random tokens
  3. per-event reduction  sum_p mask * log_prob - log w
`value` = events of all ranks / max-over-ranks device time, inputs resident in HBM.
`e2e`   = same through the public API with pinned HOST buffers: H2D of the step's inputs and D2H of the
          per-event result inside the timed region.
`parity`= the first events of the timed GPU step against the fp64 CPU oracles on the same seeded inputs; a
          run whose parity exceeds the gate prints value = null and exits 1 (a fast wrong kernel is not a result).
The JSON line also carries `roofline` (dominant kernel: the fused flow kernel), `cpu_baseline`
(CPU arm timed on the box's host cores, bounded sample) and `clocks`.
`--impl reference` times the CPU arm alone: the UNMODIFIED reference torch RAMBO from baseline/_ref (pip-installed copy
of the reference) when it is importable, else the C oracle port, plus the numpy-fp32 flow oracle (the
reference's zuko dependency is not available offline, so the flow part is always the oracle restatement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MASSES = [125.25, 172.5, 172.5, 1e-5]
SQRT_S = 13000
P_OBJ = 10
TF_A = dict(features=3, context=65, transforms=5, bins=16, hidden=[128] * 4, passes=3, bound=1.0)
METRIC = "events/s flow log_prob + PS Jacobian (ttH, 1M evts)"
WORKLOAD = ("ttH 1M events/GPU: RAMBO n=4 forward+Jacobian + TF-A transfer-flow log_prob "
            "(D=3,C=65,T=5,K=16,[128]x4) on 10 objects/event + per-event reduce")


GEMM_DESC = {"exact": "fp32 FMA on CUDA cores (flow_kernel)",
             "tc_3xf16": "tcgen05, fp16 hi/lo split operands, 3 MMAs per product, fp32 accumulate (fp32-class accuracy)",
             "tc_bf16": "tcgen05, bf16 operands, fp32 accumulate (throughput mode, ~1e-3 log_prob error)"}
KERNEL_NAME = {"exact": "flow_kernel<float,8> (fused MAF log_prob, CUDA cores)",
               "tc_3xf16": "flow_tc_kernel<3> (fused MAF log_prob, tcgen05 3xFP16)",
               "tc_bf16": "flow_tc_kernel<1> (fused MAF log_prob, tcgen05 bf16)"}


def flops_per_object():
    D, C, H, L, K, T = 3, 65, 128, 4, 16, 5
    return T * 2 * ((D + C) * H + (L - 1) * H * H + H * D * (3 * K - 1))


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop = gpu_index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def synth_inputs(events, seed, torch, pin=False):
    """Synthetic ttH-shaped batch, generated on the CPU with fixed seeds (identical for CPU and GPU legs)."""
    g = torch.Generator().manual_seed(seed)
    r = torch.rand(events, 10, generator=g).clamp_(1e-6, 1 - 1e-6)
    x = (0.35 * torch.randn(events, P_OBJ, 3, generator=g)).clamp_(-0.999, 0.999)
    ctx = torch.randn(events, P_OBJ, 65, generator=g)
    ctx[:, :, -1] = torch.arange(P_OBJ, dtype=torch.float32)  # position encoding column
    mask = torch.ones(events, P_OBJ, dtype=torch.uint8)
    mask[:, 6:] = (torch.rand(events, P_OBJ - 6, generator=g) < 0.8).to(torch.uint8)
    out = [r, x, ctx, mask]
    if pin:
        out = [t.pin_memory() for t in out]
    return out


def build_flow(torch, device, gemm="exact"):
    from memflow_msci_project_b200 import flows
    torch.manual_seed(7)
    flow = flows.NSF(features=3, context=65, transforms=5, bins=16, hidden_features=[128] * 4, randperm=False,
                     base=flows.BoxUniform, base_args=[-torch.ones(3), torch.ones(3)],
                     univariate_kwargs={"bound": 1.0}, passes=3, gemm=gemm)
    return flow.to(device)


def _import_reference_phasespace():
    """memflow.phasespace of the unmodified reference, from the in-repo install baseline/_ref (it travels to the GPU box;
    /root/reference does not).  Returns None when that install is missing."""
    import types
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "memflow")):
        return None
    for name in ("particle", "vector", "awkward"):  # imported by the reference, unused on this path
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["particle"], "Particle"):
        sys.modules["particle"].Particle = object
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        from memflow.phasespace.phasespace import PhaseSpace
        return PhaseSpace
    except Exception:
        return None


def cpu_flow_leg(spec, xn, cn, threads, dtype=np.float32):
    """numpy flow oracle on all host cores: numpy releases the GIL inside its kernels, so one Python thread per core,
    each on its own slice of the objects (BLAS limited to one thread per slice), uses the whole host."""
    from concurrent.futures import ThreadPoolExecutor
    from threadpoolctl import threadpool_limits
    from oracle import flow_oracle as fo
    bounds = np.linspace(0, xn.shape[0], threads + 1).astype(np.int64)
    with threadpool_limits(limits=1), ThreadPoolExecutor(max_workers=threads) as ex:
        def one_slice(i, chunk=8192):  # cache-sized pieces inside the slice
            a, b = int(bounds[i]), int(bounds[i + 1])
            return np.concatenate([fo.flow_log_prob(spec, xn[k:min(k + chunk, b)], cn[k:min(k + chunk, b)], dtype=dtype)
                                   for k in range(a, b, chunk)]) if b > a else np.empty(0, dtype)
        parts = list(ex.map(one_slice, range(threads)))
    return np.concatenate(parts)


def cpu_reference_leg(events, threads, seed=1234, rambo="auto"):
    """The hot path on the host cores.  RAMBO: the unmodified reference (torch, baseline/_ref) when importable
    (rambo="auto"/"reference"), else the C oracle (OpenMP); flow: numpy-fp32 oracle.  Returns a dict."""
    import torch
    from helpers import spec_from_module
    from memflow_msci_project_b200.phasespace._desc import build_rambo_desc
    from oracle import rambo as orc
    torch.set_num_threads(threads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    r, x, ctx, mask = synth_inputs(events, seed, torch)
    flow = build_flow(torch, "cpu")
    spec = spec_from_module(flow)
    xn, cn = x.reshape(-1, 3).numpy(), ctx.reshape(-1, 65).numpy()
    RefPS = _import_reference_phasespace() if rambo in ("auto", "reference") else None
    t0 = time.perf_counter()
    if RefPS is not None:
        ps = RefPS(SQRT_S, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor(MASSES), dev="cpu")
        with torch.no_grad():
            _, w, _, _ = ps.get_momenta_from_ps(r.double())   # fp64 uniforms: the SURVEY section 7 parity protocol
        w = w.numpy()
        rambo_kind = "unmodified reference torch RAMBO (baseline/_ref)"
    else:
        desc = build_rambo_desc(torch.tensor(MASSES), SQRT_S)
        _, w, _, _ = orc.forward_f32(desc, r.numpy())
        rambo_kind = "C oracle port (OpenMP)"
    t1 = time.perf_counter()
    lp = cpu_flow_leg(spec, xn, cn, threads)
    ev = (lp.reshape(events, P_OBJ) * mask.numpy()).sum(1) - np.log(w)
    t2 = time.perf_counter()
    return {"events_per_s": events / (t2 - t0), "seconds": t2 - t0, "rambo_events_per_s": events / (t1 - t0),
            "flow_objects_per_s": events * P_OBJ / (t2 - t1), "rambo_kind": rambo_kind,
            "checksum": float(np.nan_to_num(ev, neginf=0.0).sum())}


def config_dict(args, world):
    """Identical for both arms (the driver compares them)."""
    events = args.events if args.scaling == "weak" else args.events // world
    return {"workload": WORKLOAD, "events_per_gpu": events, "objects_per_event": P_OBJ,
            "rambo_compute": "f64", "flow_gemm": GEMM_DESC[args.gemm],
            "l2": "inputs (2.8 GB/step/GPU at 1M events) exceed the 126 MB L2, no flush needed"}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.ref_events
    vals = []
    for i in range(args.warmup + args.steps):
        leg = cpu_reference_leg(sample, threads, seed=1234 + i)
        if i >= args.warmup:
            vals.append(leg)
    value = float(np.mean([v["events_per_s"] for v in vals]))
    ms = float(np.mean([v["seconds"] for v in vals]) * 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world),
            "cpu_baseline": {"value": value, "unit": "events/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} events/step; RAMBO: {vals[-1]['rambo_kind']} "
                                       f"({vals[-1]['rambo_events_per_s']:.3g} events/s alone); flow: numpy fp32 oracle, one "
                                       f"slice per core ({vals[-1]['flow_objects_per_s']:.3g} objects/s; zuko is unavailable "
                                       "offline, so the flow part is the oracle restatement -> kind 'port')"},
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
PARITY_EVENTS = 10_000
PARITY_GATE = {"event_q999": 1e-5, "event_max": 1e-4, "momenta_max": 1e-5}


def parity_block(torch, host, ev_gpu, mom_gpu, flow, dev):
    """First PARITY_EVENTS events of the step the bench timed (rank 0) against the fp64 oracles on the same inputs:
    C RAMBO oracle (pinned to the reference's golden vectors) and the numpy fp64 flow oracle."""
    from helpers import momenta_rel_err, spec_from_module
    from memflow_msci_project_b200.phasespace._desc import build_rambo_desc
    from oracle import rambo as orc
    n = min(PARITY_EVENTS, host[0].shape[0])
    r, x, ctx, mask = [t[:n] for t in host]
    desc = build_rambo_desc(torch.tensor(MASSES), SQRT_S)
    mom_ref, w_ref, logw_ref, *_ = orc.forward(desc, r.double().numpy())
    spec = spec_from_module(flow)
    threads = os.cpu_count() or 1
    xn, cn = x.reshape(-1, 3).double().numpy(), ctx.reshape(-1, 65).double().numpy()
    lp_ref = cpu_flow_leg(spec, xn, cn, threads, dtype=np.float64)
    ev_ref = (lp_ref.reshape(n, P_OBJ) * mask.numpy()).sum(1) - logw_ref
    ev = ev_gpu[:n].double().cpu().numpy()
    fin = np.isfinite(ev_ref)
    same_inf = bool(np.array_equal(np.isfinite(ev), fin))
    err = np.abs(ev[fin] - ev_ref[fin]) / np.maximum(1.0, np.abs(ev_ref[fin]))
    with torch.no_grad():
        lp_gpu = flow(ctx.to(dev)).log_prob(x.to(dev)).reshape(-1).double().cpu().numpy()
    finl = np.isfinite(lp_ref)
    lerr = np.abs(lp_gpu[finl] - lp_ref[finl]) / np.maximum(1.0, np.abs(lp_ref[finl]))
    merr = momenta_rel_err(mom_gpu[:n].double().cpu().numpy(), mom_ref)
    out = {"n_events": int(n), "n_objects": int(n * P_OBJ), "against": "fp64 oracles (oracle/rambo_oracle.c, oracle/flow_oracle.py)",
           "event_log_prob_rel": {"median": float(np.median(err)), "q999": float(np.quantile(err, 0.999)), "max": float(err.max())},
           "object_log_prob_rel": {"median": float(np.median(lerr)), "q999": float(np.quantile(lerr, 0.999)),
                                   "max": float(lerr.max()), "north_star_tol": 1e-5,
                                   "frac_above_1e-5": float((lerr > 1e-5).mean())},
           "momenta_rel_max": float(merr.max()), "same_inf_pattern": same_inf, "gate": PARITY_GATE}
    out["pass"] = bool(same_inf and out["event_log_prob_rel"]["q999"] <= PARITY_GATE["event_q999"]
                       and out["event_log_prob_rel"]["max"] <= PARITY_GATE["event_max"]
                       and out["momenta_rel_max"] <= PARITY_GATE["momenta_max"])
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from memflow_msci_project_b200 import _cabi
    from memflow_msci_project_b200.likelihood import event_log_prob
    from memflow_msci_project_b200.phasespace.phasespace import PhaseSpace

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = config_dict(args, world)
    events = cfg["events_per_gpu"]
    host = synth_inputs(events, 1234 + rank, torch, pin=True)
    r, x, ctx, mask = [t.to(dev) for t in host]
    # check_inputs="deferred": the kernel records the reference's NaN / below-threshold conditions in per-event flags
    # (no device->host sync inside the step); they are checked once after the timed region
    ps = PhaseSpace(SQRT_S, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor(MASSES), dev=dev,
                    check_inputs="deferred")
    flow = build_flow(torch, dev, args.gemm)

    def step_device():
        with torch.no_grad():
            return event_log_prob(ps, flow, r, x, ctx, mask)

    from memflow_msci_project_b200.likelihood import HostPipeline
    pipe = HostPipeline(ps, flow, events, P_OBJ, 3, 65, 10, dev, chunks=args.e2e_chunks)

    def step_e2e():
        return pipe.run(*host)  # pinned host inputs -> chunked H2D overlapped with the kernels -> pinned host result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # per-kernel timing of the dominant kernel (fused flow log_prob), CUDA events on the launch stream
    def flow_only():
        with torch.no_grad():
            return flow(ctx).log_prob(x)

    step_device()                      # first call packs the flow weights (not part of a step)
    n0 = _cabi.launch_count()
    ev_gpu, mom_gpu, _ = step_device()
    launches_per_step = _cabi.launch_count() - n0   # RAMBO forward + fused flow log_prob + per-event reduce

    with ClockSampler(local) as clk:
        ms_total = timed(step_device, args.steps, args.warmup)
        ms_flow = timed(flow_only, max(2, args.steps // 2), 1) / max(2, args.steps // 2)
        ms_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    clocks = clk.summary()
    ps.generator.raise_if_flagged()   # the deferred input checks of every step above
    ms_step = ms_total / args.steps
    value = world * events / (ms_step * 1e-3)
    e2e_value = world * events / (ms_e2e / args.steps * 1e-3)

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md sustained 1.4 PFLOP/s)"
    achieved_tf = flops_per_object() * events * P_OBJ / (ms_flow * 1e-3) / 1e12
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "flow_kernel_traffic.json")
    if os.path.exists(tfile):
        traffic = json.load(open(tfile)).get("dram_bytes_per_object", 0) * events * P_OBJ or None

    rc = 0
    if rank == 0:
        parity = None if args.no_parity else parity_block(torch, host, ev_gpu, mom_gpu, flow, dev)
        h2d = sum(t.numel() * t.element_size() for t in host)
        line = {"metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "gpu_launches": int(launches_per_step * args.steps),
                "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(events * 4),
                        "how": f"HostPipeline: {args.e2e_chunks} chunks, H2D of chunk i+1 overlaps the kernels of chunk i; "
                               "the host blocks on the D2H copy of the per-event result every step"},
                "roofline": {"bound": "tensor", "kernel": KERNEL_NAME[args.gemm],
                             "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                             "traffic": traffic, "peak_source": peak_src, "kernel_ms": ms_flow,
                             "algorithmic_flops_per_object": flops_per_object(),
                             "note": "achieved = ALGORITHMIC dense FLOPs / kernel time; the 3xFP16 mode issues 3 MMAs per "
                                     "algorithmic product, so tensor-pipe work is 3x the algorithmic figure"},
                "clocks": clocks}
        if parity is not None:
            line["parity"] = parity
            if not parity["pass"]:
                line["value"] = None
                line["parity_failed"] = "GPU step deviates from the fp64 oracles beyond the gate: no throughput is reported"
                rc = 1
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            leg = cpu_reference_leg(args.ref_events, threads)
            line["cpu_baseline"] = {"value": leg["events_per_s"], "unit": "events/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.ref_events} events, {leg['seconds']:.1f} s; RAMBO: {leg['rambo_kind']} "
                                              f"({leg['rambo_events_per_s']:.3g} events/s alone); flow: numpy fp32 oracle, one slice "
                                              f"per core ({leg['flow_objects_per_s']:.3g} objects/s; zuko unavailable offline -> 'port')"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=int, default=1_000_000,
                    help="events per GPU per step (--scaling weak) or in total (--scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: `events` per GPU (the headline: 1M events per GPU); strong: `events` split over the ranks")
    ap.add_argument("--ref-events", type=int, default=600_000, help="events per step of the CPU arm (about 10 s on the bench box, 30 s on 8 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed step (profiling runs)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="chunks of the host->device pipeline of the e2e leg")
    ap.add_argument("--gemm", default="tc_3xf16", choices=["exact", "tc_3xf16", "tc_bf16"],
                    help="conditioner GEMM arithmetic of the flow kernel (default: tensor cores at fp32-class accuracy)")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
