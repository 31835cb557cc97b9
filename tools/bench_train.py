#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[3]): transfer-flow likelihood + sampling loss, batch 64k events
per GPU (x10 objects), Adam, and ONE flat NCCL gradient all-reduce per step under torchrun.

The forward of log_prob and rsample runs in the fused sm_100a kernels; the backward recomputes the flow transform by
transform (flows/_autograd.py): hyper-network GEMMs (recompute, dX, dW) on tcgen05 through
mf_gemm_nt_3xf16 (MEMFLOW_B200_TORCH_GEMM=1: torch / cuBLAS, the round-1 backward), spline forward / backward in the
native kernels mf_rqs_transform / mf_rqs_backward.

    python tools/bench_train.py [--events 65536] [--steps 3] [--warmup 2] [--no-sampling-loss]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py ...
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from memflow_msci_project_b200 import _cabi, flows
from memflow_msci_project_b200.sharding import all_reduce_gradients


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=int, default=65536, help="events per GPU")
    ap.add_argument("--objects", type=int, default=10)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-sampling-loss", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(3)  # same initial weights on every rank
    flow = flows.NSF(features=3, context=65, transforms=5, bins=16, hidden_features=[128] * 4, passes=3,
                     base=flows.BoxUniform, base_args=[-torch.ones(3), torch.ones(3)], univariate_kwargs={"bound": 1.0}).to(dev)
    opt = torch.optim.Adam(flow.parameters(), lr=1e-4)
    n = args.events * args.objects
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = (0.3 * torch.randn(n, 3, generator=g, device=dev)).clamp_(-0.99, 0.99)
    ctx = torch.randn(n, 65, generator=g, device=dev)
    mask = (torch.rand(args.events, args.objects, generator=g, device=dev) < 0.8).float()
    nparam = sum(p.numel() for p in flow.parameters())

    def step():
        opt.zero_grad(set_to_none=True)
        dist_c = flow(ctx)
        logp = dist_c.log_prob(x).view(args.events, args.objects)            # transfer_flow_paper_AllPartons_btag.py:202-216
        loss = -(logp * mask).sum(1).mean()
        if not args.no_sampling_loss:
            xs = dist_c.rsample()                                              # sampling loss (Huber), run_model_labframe_sampling.py:466-496
            loss = loss + torch.nn.functional.huber_loss(xs, x)
        loss.backward()
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            all_reduce_gradients(flow.parameters())
            e1.record()
            ar_events.append((e0, e1))
        opt.step()
        return loss

    ar_events = []

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _cabi.launch_count()
    ar_events.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ar_ms = sum(e0.elapsed_time(e1) for e0, e1 in ar_events) / max(1, args.steps)   # flat-buffer gradient all-reduce per step
    if rank == 0:
        print(json.dumps({"config": "transfer-flow training step: likelihood" + ("" if args.no_sampling_loss else " + sampling (Huber) loss") +
                          ", TF-A (D=3,C=65,T=5,K=16,[128]x4), Adam, flat NCCL gradient all-reduce",
                          "events_per_gpu": args.events, "objects_per_event": args.objects, "n_gpus": world, "ms_per_step": ms,
                          "events_per_s": args.events * world / ms * 1e3, "all_reduce_ms_per_step": ar_ms, "parameters": nparam, "loss": float(loss.item()),
                          "fused_kernel_launches_per_step": (_cabi.launch_count() - l0) / args.steps,
                          "backward": ("recompute per transform: torch hyper-network (cuBLAS)" if os.environ.get("MEMFLOW_B200_TORCH_GEMM") == "1" else "recompute per transform: hyper-network GEMMs on tcgen05 (mf_gemm_nt_3xf16)") + " + native spline forward/backward kernels", "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
