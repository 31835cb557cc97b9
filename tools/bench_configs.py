#!/usr/bin/env python
"""Secondary benchmarks: BASELINE.json configs 1, 3 and 5 on one GPU (or sharded under torchrun), next to the
headline bench.py (config 2).  Device-resident inputs, CUDA events, max over ranks; prints one JSON line per config.

    python tools/bench_configs.py [--events N] [--configs 1,3,5]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_configs.py ...
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from memflow_msci_project_b200 import flows
from memflow_msci_project_b200.phasespace.phasespace import PhaseSpace
from memflow_msci_project_b200.sharding import shard_range
from memflow_msci_project_b200.unfolding_flow.utils import Compute_ParticlesTensor

MASSES4 = [125.25, 172.5, 172.5, 1e-5]
MASSES8 = [4.7, 4.7, 4.7, 1e-3, 1e-3, 4.7, 1e-3, 1e-3]


def uf_flow(dev, gemm):
    """configs/flow_pretrain_labframe_sampling/flow_labframe_sampling_v3f.yaml:80-94"""
    torch.manual_seed(11)
    return flows.NSF(features=10, context=16, transforms=6, bins=40, hidden_features=[512] * 2, randperm=True,
                     base=flows.DiagNormal, base_args=[torch.zeros(10), 0.35 * torch.ones(10)],
                     univariate_kwargs={"bound": 1.1}, passes=2, gemm=gemm).to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,5")
    ap.add_argument("--events3", type=int, default=1_000_000, help="config 3: total events (sharded over ranks)")
    ap.add_argument("--events5", type=int, default=1_000, help="config 5: events (x points each), sharded over ranks")
    ap.add_argument("--points5", type=int, default=1_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--gemm", default="mma_3xf16", choices=["exact", "mma_3xf16"],
                    help="conditioner arithmetic of the width-512 unfolding flow (tcgen05 modes stop at width 128)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def timed(fn):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    cfgs = [int(c) for c in args.configs.split(",")]
    with torch.no_grad():
        if 1 in cfgs:
            # config 1: RAMBO map + Jacobian, semileptonic ttH (8 partons); 1e5 events is the reference's CPU case,
            # timed here at 8e6 events so that the tensors exceed L2
            for name, masses in (("n8", MASSES8), ("n4", MASSES4)):
                n = len(masses)
                for compute in ("f64", "f32"):
                    ps = PhaseSpace(13000, [21, 21], [0] * n, final_masses=torch.tensor(masses), dev=dev, compute=compute, check_inputs=False)
                    N = 8_000_000
                    r = torch.rand(N, 3 * n - 2, device=dev).clamp_(1e-6, 1 - 1e-6)
                    ms = timed(lambda: ps.get_momenta_from_ps(r))
                    emit(config=1, what=f"RAMBO forward + Jacobian {name}, fp32 I/O, compute={compute}", events=N, ms=ms,
                         events_per_s=N / ms * 1e3, n_gpus=world)
                    del r
        if 3 in cfgs:
            # config 3: unfolding flow sample -> sigmoid -> RAMBO forward -> get_PS (RAMBO inverse + mask) -> log_prob
            a, b = shard_range(args.events3, rank, world)
            N = b - a
            flow = uf_flow(dev, args.gemm)
            ps = PhaseSpace(13000, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor(MASSES4), dev=dev, check_inputs=False)
            g = torch.Generator(device=dev).manual_seed(11 + rank)
            noise = torch.randn(N, 10, generator=g, device=dev)
            ctx = torch.randn(N, 16, generator=g, device=dev)

            from memflow_msci_project_b200 import flows as _fl
            sig, lgt = _fl.SigmoidAffine(1.0, 0.0), _fl.LogitAffine(0.0, 1.0, eps=5e-5)

            def step3():
                # scripts/run_model_labframe_sampling.py:426-427: sample + sigmoid bridge fused into the sampling kernel
                u = flow(ctx).sample_from_noise(noise, post=sig)
                mom, w, x1, x2 = ps.get_momenta_from_ps(u)                    # :432
                boost = torch.stack(((x1 + x2) * 6500.0, torch.zeros_like(x1), torch.zeros_like(x1), (x1 - x2) * 6500.0), 1)
                back, _, mask = Compute_ParticlesTensor.get_PS(mom[:, 2:].contiguous(), boost)  # unfolding_flow.py:163
                lp = flow(ctx).log_prob(back, pre=lgt)   # :170-179: logit(eps=5e-5) bridge fused into the log_prob kernel
                return lp, mask

            ms = timed(step3)
            emit(config=3, what=f"UF (D=10,C=16,T=6,K=40,[512]x2) sample + RAMBO n=4 forward + get_PS + log_prob, flow gemm={args.gemm}",
                 events_total=args.events3, events_per_gpu=N, ms=ms, events_per_s=args.events3 / ms * 1e3, n_gpus=world)
        if 5 in cfgs:
            # config 5: MEM sweep, importance-sampled points per event; context hoisted per event (ctx_group)
            a, b = shard_range(args.events5, rank, world)
            E, P = b - a, args.points5
            flow = uf_flow(dev, args.gemm)
            ps = PhaseSpace(13000, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor(MASSES4), dev=dev, check_inputs=False)
            g = torch.Generator(device=dev).manual_seed(5 + rank)
            r = torch.rand(E * P, 10, generator=g, device=dev).clamp_(1e-4, 1 - 1e-4)
            ctx = torch.randn(E, 16, generator=g, device=dev)

            def step5():
                mom, w, x1, x2, logw = ps.generator.generateKinematics_batch(r, want_logw=True)
                lp = flow(ctx, ctx_group=P).log_prob(torch.logit(r))   # one context row per event, read in place
                return lp - logw                                        # notebooks/TestFlows_ImportanceSampling.ipynb cell 52

            ms = timed(step5)
            emit(config=5, what=f"MEM sweep: RAMBO n=4 forward + UF log_prob, context hoisted per event, flow gemm={args.gemm}", events=args.events5,
                 points_per_event=P, ms=ms, points_per_s=args.events5 * P / ms * 1e3, n_gpus=world)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
