"""CPU: the N>1 host logic on the gloo backend, world_size 2 (no GPU needed): contiguous event
sharding tiles the batch, gathering restores it, and the flat gradient all-reduce equals the per-tensor one."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from memflow_msci_project_b200.sharding import all_reduce_gradients, gather_events, shard, shard_range


@pytest.mark.parametrize("n,world", [(10, 1), (10, 2), (10, 3), (7, 8), (1_000_000, 8), (0, 4)])
def test_shard_ranges_tile_the_batch(n, world):
    ranges = [shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        full = torch.arange(1001 * 3, dtype=torch.float32).reshape(1001, 3)
        mine = shard(full)
        result = mine * 2 + rank * 0            # stand-in for the per-event computation
        back = gather_events(result, full.shape[0])
        ok_gather = torch.equal(back, full * 2)
        # gradient all-reduce: flat buffer vs one all_reduce per tensor
        model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2))
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(16, 4, generator=g)
        model(x).pow(2).sum().backward()
        ref = []
        for p in model.parameters():
            t = p.grad.clone()
            dist.all_reduce(t)
            ref.append(t / world)
        n = all_reduce_gradients(model.parameters())
        ok_grad = all(torch.allclose(p.grad, r, atol=1e-6) for p, r in zip(model.parameters(), ref))
        out[rank] = (ok_gather, ok_grad, n, tuple(mine.shape))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_sharding_and_flat_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        ok_gather, ok_grad, n, shape = out[r]
        assert ok_gather and ok_grad and n == 4 * 8 + 8 + 8 * 2 + 2
    assert out[0][3][0] + out[1][3][0] == 1001
