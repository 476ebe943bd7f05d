"""world_size-2 gloo test of the data-parallel host logic (runs on CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from safevla_b200.parallel import TAIL, allreduce_arena, shard_samplers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1000
    g = torch.Generator().manual_seed(rank)
    local = torch.randn(n, generator=g)
    comm = torch.zeros(n + TAIL)
    comm[:n] = local
    cost = torch.tensor([3.0 + rank, 2.0])
    scale = allreduce_arena(comm, n, cost)
    lo, hi = shard_samplers(64, world, rank)
    out[rank] = (comm.clone(), scale, lo, hi)
    # K = 2 cost channels: two (sum, count) pairs ride in the same tail
    comm[:n] = local
    allreduce_arena(comm, n, torch.tensor([1.0 + rank, 2.0, 10.0 * (rank + 1), 2.0]))
    out[10 + rank] = comm[n:n + 4].clone()
    # a non-final repeat zeroes the tail so stale sums never leak into lambda
    comm[:n] = local
    allreduce_arena(comm, n, None)
    assert comm[n:].abs().sum() == 0
    dist.destroy_process_group()


def test_single_allreduce_carries_grads_and_cost_scalars():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    g0, g1 = torch.Generator().manual_seed(0), torch.Generator().manual_seed(1)
    exp = torch.randn(1000, generator=g0) + torch.randn(1000, generator=g1)
    for r in range(world):
        comm, scale, lo, hi = out[r]
        assert torch.allclose(comm[:1000], exp) and scale == 0.5
        assert comm[1000].item() == 7.0 and comm[1001].item() == 4.0  # sum of costs, episode count
        assert (lo, hi) == (r * 32, (r + 1) * 32)
    assert torch.equal(out[0][0], out[1][0])  # identical on every rank -> identical lambda without a broadcast
    assert out[10].tolist() == out[11].tolist() == [3.0, 4.0, 30.0, 4.0]


def test_shard_samplers_rejects_ragged():
    import pytest
    with pytest.raises(ValueError):
        shard_samplers(10, 4, 0)
