"""Launched by tests/test_multigpu_gpu.py under torchrun (2 ranks, NCCL): one view-sharded training step, then the same two
shards run one after the other on this rank alone.  The all-reduced gradients of the replicated parameters must equal the
mean of the two serial shards (the plumbing of gshell_b200/distributed.py on hardware: flat-bucket mean all-reduce, per-rank
RNG streams).  The two loss terms that are not sums over views (shading_loss's ratio of batch means, the close-mSDF
regulariser's union of visible triangles) are covered on gloo by tests/test_distributed_cpu.py and switched off here."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if os.environ.get("GSB_HOST_EMULATION") == "1":
        # no GPU: the host build of the kernels (tests/conftest.py::bind_host_library) on CPU tensors, gloo for the collectives --
        # the same step, sharding, exchange and reduction code (tests/test_emulated_gpu_suite_cpu.py)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import conftest
        conftest.bind_host_library()
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
    from gshell_b200 import synthetic
    from gshell_b200.denoiser.denoiser import BilateralDenoiser
    from gshell_b200.distributed import allreduce_mean_grads_, shard_views
    from gshell_b200.geometry.gshell_tets_geometry import GShellTetsGeometry, default_flags
    from gshell_b200.grids import save_tets_npz
    from gshell_b200.render import light
    from gshell_b200.render import render as _render
    from gshell_b200.render import renderutils as ru
    from gshell_b200.render.optixutils import ops as _ops
    n_views, res, n = 4, ([64, 64] if os.environ.get("GSB_HOST_EMULATION") == "1" else [96, 96]), 3      # rows must split into 8-row blocks per rank
    npz = os.path.join(tempfile.gettempdir(), f"gsb_two_rank_{rank}.npz")
    save_tets_npz(npz, 10)
    FLAGS = default_flags(n_samples=n, sphere_init=True, lambda_diffuse=0.0, lambda_specular=0.0, msdf_reg_close_scale=0.0)
    rng = np.random.RandomState(7)
    gen = torch.Generator().manual_seed(7)
    mvp_all, campos_all = synthetic.random_cameras(n_views, res, "cpu", rng)
    img_all, bg_all = synthetic.random_target(n_views, res, "cpu", gen)
    loss_fn = lambda a, b: ru.image_loss(a, b, loss="l1", tonemapper="log_srgb")    # noqa: E731

    def shard_grads(r, seed=None):
        """Gradients of (sdf, msdf, deform, light) from the views of rank r, with rank r's RNG streams."""
        torch.manual_seed(0)
        geo = GShellTetsGeometry(64, 2.0, FLAGS, tet_init_file=npz, device=dev)
        lgt = light.create_trainable_env_rnd(16, scale=0.5, bias=0.25, device=dev)
        idx = list(shard_views(n_views, r, 2))
        mat = synthetic.LeafMaterialField(len(idx), res[0], res[1], dev, torch.Generator().manual_seed(100 + r))
        target = {"mvp": mvp_all[idx].to(dev), "campos": campos_all[idx].to(dev), "img": img_all[idx].to(dev),
                  "background": bg_all[idx].to(dev), "resolution": res, "spp": 1}
        torch.manual_seed(1 + r)
        _render.rnd_seed = r * 1000003 if seed is None else seed
        lgt.update_pdf()
        il, dl, rl = geo.tick(None, target, lgt, {"kd_ks": mat, "bsdf": "pbr"}, loss_fn, 1200, BilateralDenoiser())
        (il + dl + rl).backward()
        return [geo.sdf, geo.msdf, geo.deform, lgt.base]

    # (1) plumbing, with every pixel shaded at home: the sharded step equals the two shards run one after the other
    _ops.BALANCE_SHADING = False
    params = shard_grads(rank)
    allreduce_mean_grads_(params)
    got = [p.grad.clone() for p in params]
    def serial_mean():
        serial = [shard_grads(r) for r in range(2)]
        return [(a.grad + b.grad) / 2 for a, b in zip(*serial)]
    want = serial_mean()
    again = serial_mean()
    ok = True
    for name, g, w, w2 in zip(("sdf", "msdf", "deform", "light"), got, want, again):
        err = float((g - w).abs().max())
        scale = float(w.abs().max())
        # the kernels accumulate with float atomics (light gradient, vertex scatter): two serial evaluations of the SAME thing
        # differ by that much, so the bar is 1e-5 relative or a few times this run-to-run noise, whichever is larger
        noise = float((w - w2).abs().max())
        print(f"rank {rank} {name}: max err {err:.3e} scale {scale:.3e} serial run-to-run noise {noise:.3e}", file=sys.stderr)
        ok &= err <= max(1e-5 * scale, 4.0 * noise) and scale > 0
    # (2) the row-block exchange of the forward shading pass (ops.BALANCE_SHADING): same samples per pixel -- ids rank * B*H*W + i,
    # one seed -- shaded at home or dealt out over the ranks give the same gradients
    npix = (n_views // 2) * res[0] * res[1]

    def step(balanced):
        _ops.BALANCE_SHADING = balanced
        _ops.PIXEL_ID_BASE = None if balanced else rank * npix
        ps = shard_grads(rank, seed=77)
        allreduce_mean_grads_(ps)
        _ops.PIXEL_ID_BASE = None
        return [p.grad.clone() for p in ps]
    dealt, home, home2 = step(True), step(False), step(False)
    for name, g, w, w2 in zip(("sdf", "msdf", "deform", "light"), dealt, home, home2):
        err = float((g - w).abs().max())
        scale = float(w.abs().max())
        noise = float((w - w2).abs().max())
        print(f"rank {rank} dealt-vs-home {name}: max err {err:.3e} scale {scale:.3e} run-to-run noise {noise:.3e}", file=sys.stderr)
        ok &= err <= max(1e-5 * scale, 4.0 * noise) and scale > 0
    # both ranks hold the same reduced gradients
    flat = torch.cat([g.reshape(-1) for g in got])
    other = flat.clone()
    dist.broadcast(other, src=0)
    ok &= bool(torch.equal(flat, other))
    os.unlink(npz)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if float(t) != 1.0:
        sys.exit(3)
    if rank == 0:
        print("TWO_RANK_OK")


if __name__ == "__main__":
    main()
