"""N>1 host logic on CPU: world_size-2 gloo all-reduce of the flat gradient bucket and the view sharding."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gshell_b200.distributed import allreduce_mean_grads_, shard_views


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.zeros(5, 3))
    b = torch.nn.Parameter(torch.zeros(7))
    c = torch.nn.Parameter(torch.zeros(2))          # no grad on rank 1
    a.grad = torch.full((5, 3), float(rank + 1))
    b.grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
    if rank == 0:
        c.grad = torch.ones(2) * 4
    allreduce_mean_grads_([a, b, c])
    out[rank] = (a.grad.clone(), b.grad.clone(), c.grad.clone())
    dist.destroy_process_group()


def test_flat_bucket_allreduce_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29511, out), nprocs=world, join=True)
    for rank in range(world):
        a, b, c = out[rank]
        assert torch.allclose(a, torch.full((5, 3), 1.5))
        assert torch.allclose(b, torch.arange(7, dtype=torch.float32) * 1.5)
        assert torch.allclose(c, torch.ones(2) * 2)


def test_shard_views():
    assert [list(shard_views(8, r, 4)) for r in range(4)] == [[0, 1], [2, 3], [4, 5], [6, 7]]
    got = [list(shard_views(10, r, 4)) for r in range(4)]
    assert sum(got, []) == list(range(10)) and [len(g) for g in got] == [3, 3, 2, 2]
