"""Multi-GPU plumbing: views shard across ranks, the mesh is extracted redundantly (deterministic, identical on every
rank), and ONE all-reduce per step averages the gradients of the replicated parameters (sdf | msdf | deform | light) as a
single flat bucket.  NCCL over NVLink/NVSwitch on the GPUs; the same code runs on gloo for the CPU tests."""
import torch
import torch.distributed as dist


def allreduce_mean_grads_(params, group=None):
    """In-place mean of `.grad` over all ranks through one flat bucket.  Params without a grad contribute zeros (a rank
    whose views miss the object still takes part in the collective)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= world
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n


def allreduce_or_mask_(mask, group=None):
    """In-place logical OR of a bool mask over all ranks (one small MAX all-reduce).  The close-mSDF regulariser selects the
    boundary vertices seen by ANY view of the batch (reference gshell_tets_geometry.py:343-356); with views sharded over
    ranks the union needs this exchange to stay exact."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mask
    m = mask.to(torch.uint8)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    mask.copy_(m.bool())
    return mask


def batch_mean(x, group=None):
    """Mean of `x` over the pixels of ALL ranks' views (equal view counts per rank), differentiable through the local part.
    The value is the global mean; the gradient is that of the local mean, so that the mean all-reduce of the parameter
    gradients reproduces d f(global mean) exactly for terms that are not sums over views, e.g. the
    mean(specular)/mean(diffuse) ratio of shading_loss (reference regularizer.py:39)."""
    local = x.mean()
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    tot = local.detach().clone()
    dist.all_reduce(tot, group=group)
    return local + (tot / dist.get_world_size(group) - local.detach())


def shard_views(n_views_total, rank, world):
    """Contiguous block of view indices owned by `rank` (uneven totals give the first ranks one extra view)."""
    base, extra = divmod(n_views_total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))
