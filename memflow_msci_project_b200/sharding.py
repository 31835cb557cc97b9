"""Event-parallel sharding across the GPUs of one box.

Events are independent on every piece of the path, so the multi-GPU form is: one process per GPU
(`torchrun`), each rank owns a contiguous event range, flow weights are replicated, and NOTHING is
exchanged on the data path.  The only collective of the whole path is the gradient all-reduce of the
training step (BASELINE config 4), done on ONE flat buffer so that it is latency- not launch-bound
over NVLink/NVSwitch (the reference relies on DDP's bucketed all-reduce,
scripts/run_model_labframe_sampling.py:262-266).
"""
import torch
import torch.distributed as dist


def shard_range(n_events, rank, world_size):
    """Contiguous [start, stop) of rank's events; sizes differ by at most one, all ranges tile [0, n)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside 0..{world_size - 1}")
    base, rem = divmod(int(n_events), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard(tensor, rank=None, world_size=None):
    """This rank's contiguous slice along dim 0 (a view, no copy)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    a, b = shard_range(tensor.shape[0], rank, world_size)
    return tensor[a:b]


def gather_events(local, n_events, group=None):
    """Inverse of `shard` for a per-event result tensor: every rank gets the full [n_events, ...] tensor
    (outside any timed region; the path itself never needs it)."""
    world = dist.get_world_size(group)
    sizes = [b - a for a, b in (shard_range(n_events, r, world) for r in range(world))]
    width = max(sizes)  # all_gather wants equal shapes: pad the short shards by one row
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    outs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(outs, padded, group=group)
    return torch.cat([o[:k] for o, k in zip(outs, sizes)], dim=0)


def all_reduce_gradients(parameters, group=None, average=True):
    """Sum (or average) the gradients of `parameters` over the ranks with ONE all-reduce of a flat buffer."""
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()
