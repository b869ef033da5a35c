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


def _worker_exact(rank, world, port, out):
    """Two ranks shade two 'views' each; the shading_loss ratio term and the visibility union must give the gradient of the
    single-process batch of four views after the mean all-reduce (SURVEY 8e exactness caveats)."""
    from gshell_b200.distributed import allreduce_or_mask_, batch_mean
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    theta0 = torch.rand(3, generator=g)
    views = torch.rand(4, 6, 5, 3, generator=g)                      # all four views, identical on both ranks
    theta = torch.nn.Parameter(theta0.clone())
    mine = views[2 * rank:2 * rank + 2]
    spec, diff = (mine * theta).sum(-1), (mine * theta * theta).sum(-1) + 1.0
    loss = batch_mean(spec) / batch_mean(diff) + (mine * theta).mean()
    loss.backward()
    allreduce_mean_grads_([theta])
    mask = torch.zeros(9, dtype=torch.bool)
    mask[[1, 4] if rank == 0 else [4, 7]] = True
    allreduce_or_mask_(mask)
    out[rank] = (theta.grad.clone(), mask.clone(), float(batch_mean(spec)))
    dist.destroy_process_group()


def test_batch_terms_are_exact_under_view_sharding():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_exact, args=(world, 29517, out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    theta = torch.rand(3, generator=g).requires_grad_()
    views = torch.rand(4, 6, 5, 3, generator=g)
    spec, diff = (views * theta).sum(-1), (views * theta * theta).sum(-1) + 1.0
    (spec.mean() / diff.mean() + (views * theta).mean()).backward()
    for rank in range(world):
        grad, mask, m = out[rank]
        assert torch.allclose(grad, theta.grad, rtol=1e-5, atol=1e-7)
        assert mask.nonzero().flatten().tolist() == [1, 4, 7]
        assert abs(m - float(spec.mean())) < 1e-6


def _worker_rows(rank, world, port, out):
    """Row-block exchange of the balanced shading pass: every pixel travels to the rank that shades it and back to its place."""
    from gshell_b200.render.optixutils import ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, H, W = 2, 32, 5
    npix = B * H * W
    ids = torch.arange(rank * npix, (rank + 1) * npix, dtype=torch.float32).view(B, H, W, 1)
    rows = torch.arange(H, dtype=torch.float32).view(1, H, 1, 1).expand(B, H, W, 1)
    away = ops._exchange_out(torch.cat([ids, rows], -1), world)          # [world * B, H / world, W, 2]
    # this rank received row blocks rank, rank + world, ... of the images of EVERY rank
    got_rows = away[..., 1]
    want_rows = torch.cat([torch.arange(8) + 8 * (rank + world * j) for j in range(H // (8 * world))]).float()
    ok = bool((got_rows == want_rows.view(1, -1, 1)).all())
    owners = (away[..., 0] // npix).long()                                 # source rank of every received pixel
    ok &= bool((owners.view(world, B, -1) == torch.arange(world).view(world, 1, 1)).all())
    home = ops._exchange_home(torch.cat([away[..., 0:1] * 2 + 1, away[..., 1:2]], -1), world, H)
    ok &= bool(torch.equal(home[..., 0:1], ids * 2 + 1)) and bool(torch.equal(home[..., 1:2], rows.contiguous()))
    # the exchange is only used when every rank holds the same number of views and the rows split into whole blocks
    cpu = torch.device("cpu")
    ok &= ops._balance_world(H, B, cpu) == world and ops._balance_world(H + 8, B, cpu) == 1
    ok &= ops._balance_world(H, B + rank, cpu) == 1
    out[rank] = ok
    dist.destroy_process_group()


def test_row_block_exchange_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_rows, args=(world, 29513, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_row_dealing_is_a_permutation():
    """_deal_rows hands row block j of every image to part j mod world; _collect_rows undoes it (pure index shuffles, no process group)."""
    from gshell_b200.render.optixutils import ops
    for world, B, H, W, C in ((2, 1, 16, 3, 2), (4, 2, 64, 5, 3), (8, 1, 128, 2, 1)):
        x = torch.arange(B * H * W * C, dtype=torch.float32).view(B, H, W, C)
        parts = ops._deal_rows(x, world)
        assert parts.shape == (world, B, H // world, W, C)
        for r in range(world):
            rows = torch.cat([torch.arange(8) + 8 * (r + world * j) for j in range(H // (8 * world))])
            assert torch.equal(parts[r], x[:, rows])
        assert torch.equal(ops._collect_rows(parts.contiguous(), H), x)


def test_balancing_needs_whole_row_blocks():
    from gshell_b200.render.optixutils import ops
    assert ops._balance_world(1024) == 1          # no process group: every pixel is shaded at home
