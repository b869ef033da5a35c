"""Cases added in the session that had no GPU minutes left (session 3 of round 2): the one-launch inference of the material
field, render_mesh with multisampling, tick() with the mSDF field MLP.  They exercise code that has run on the host build of the
kernels only (tests/test_emulated_gpu_suite_cpu.py), so they live in a file that sorts after the established suites: a surprise on
the device must not stop those under `-x`."""
import numpy as np
import pytest
import torch

from gshell_b200.render import mlptexture

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)
from test_render_fused_gpu import _scene

pytestmark = pytest.mark.gpu
D = DEVICE


def test_tick_with_the_msdf_field_mlp(tmp_path):
    """FLAGS.use_msdf_mlp (reference gshell_tets_geometry.py:118-136, 199-202) through the shared tick smoke of tests/test_pipeline_gpu.py"""
    from test_pipeline_gpu import test_geometry_tick_runs_and_optimises
    test_geometry_tick_runs_and_optimises("tets_msdf_mlp", tmp_path)


@pytest.mark.parametrize("channels,seed", [(6, 0), (9, 1), (3, 2)])
def test_fused_field_inference_matches_the_autograd_path(channels, seed):
    """MLPTexture3D.sample under torch.no_grad() runs the one-launch kernel (encoding -> MLP -> range map, csrc/hashgrid.cu::
    k_field_infer); with autograd it is the encoding kernel + the PyTorch MLP of the reference.  Same parameters, same points
    (some outside the box: clamped), the two must agree to fp32 rounding of the 32-term dot products."""
    dev = device()
    torch.manual_seed(seed)
    aabb = torch.tensor([[-1.0, -0.8, -1.2], [1.1, 1.0, 0.9]], device=dev)
    mn = torch.linspace(-0.2, 0.3, channels, device=dev)
    mx = mn + torch.linspace(0.5, 1.5, channels, device=dev)
    tex = mlptexture.MLPTexture3D(aabb, channels=channels, min_max=[mn, mx])
    with torch.no_grad():
        tex.encoder.params.mul_(3000.0)                 # U(-0.3, 0.3): activations of order one, both ReLU branches taken
    assert tex._fused_inference_ok()
    pos = torch.rand(2, 37, 29, 3, device=dev) * 2.6 - 1.3
    want = tex.sample(pos.clone().requires_grad_()).detach()
    with torch.no_grad():
        got = tex.sample(pos)
    assert got.shape == want.shape == (2, 37, 29, channels)
    assert bool((got >= mn - 1e-6).all()) and bool((got <= mx + 1e-6).all())
    assert float((got - want).abs().max()) <= 2e-5 * float((mx - mn).max()), float((got - want).abs().max())
    assert float(want.std()) > 1e-3                     # not a constant field
    # a configuration the kernel does not cover keeps the composed path (no error, same interface)
    wide = mlptexture.MLPTexture3D(aabb, channels=channels, internal_dims=64, min_max=[mn, mx])
    assert not wide._fused_inference_ok()
    with torch.no_grad():
        assert wide.sample(pos).shape == (2, 37, 29, channels)


@pytest.mark.parametrize("msaa", [True, False])
def test_render_mesh_multisampled(msaa):
    """render_mesh(spp = 2) (reference render.py:224-233, 403-433): visibility at 2x the resolution; with msaa the shading runs at
    `resolution`, is replicated and laid over the background at the visibility resolution, then box-filtered down.  Checked on what
    does not depend on the Monte-Carlo samples: the coverage channel of `shaded` and of `kd` equals the box-filtered antialiased
    coverage of the 2x rasterisation, every buffer has the reference's shape, gradients reach the mesh."""
    from gshell_b200 import synthetic
    from gshell_b200.geometry.gshell_tets_geometry import default_flags
    from gshell_b200.render import light, optixutils as ou, raster, render, renderutils as ru, util
    m, msdf, _, _, _, _ = _scene(seed=9)
    B, H, W, spp = 2, 40, 48, 2
    mvp, campos = synthetic.random_cameras(B, (H, W), D, np.random.RandomState(9))
    Hs, Ws = (H, W) if msaa else (H * spp, W * spp)                   # shading resolution
    field = synthetic.LeafMaterialField(B, Hs, Ws, D)
    m.material = {"kd_ks": field, "bsdf": "pbr"}
    from gshell_b200.render import mesh
    v_pos = m.v_pos.clone().requires_grad_()
    mm = mesh.auto_normals(mesh.Mesh(v_pos, m.t_pos_idx, material=m.material))
    lgt = light.create_trainable_env_rnd(16, device=D)
    bg = torch.rand(B, H, W, 3, device=D)
    bufs = render.render_mesh(default_flags(n_samples=2), None, mm, mvp, campos, lgt, [H, W], spp=spp, msaa=msaa, background=bg,
                              optix_ctx=ou.OptiXContext(), shadow_scale=0.0, use_uv=False, extra_dict={"msdf": msdf})
    for k, v in bufs.items():
        if k != "visible_triangles":
            assert v.shape == (B, H, W, 1 if k == "msdf_image" else 4) and bool(torch.isfinite(v).all()), k
    clip = ru.xfm_points(mm.v_pos[None].detach(), mvp)
    rast, _ = raster.rasterize(clip, mm.t_pos_idx.int(), (H * spp, W * spp))
    cov = (rast[..., 3:4] > 0).float()
    want_alpha = util.avg_pool_nhwc(raster.antialias(cov.contiguous(), rast, clip, mm.t_pos_idx.int()), spp)
    assert float((bufs["shaded"][..., 3:4] - want_alpha).abs().max()) < 1e-5
    assert float((bufs["kd"][..., 3:4] - want_alpha).abs().max()) < 1e-5
    frac = (want_alpha > 0) & (want_alpha < 1)
    assert int(frac.sum()) > 20                                       # the box filter alone makes the outline fractional
    (bufs["shaded"][..., 0:3].sum() + bufs["shaded"][..., 3].sum()).backward()
    assert v_pos.grad is not None and float(v_pos.grad.abs().sum()) > 0 and bool(torch.isfinite(v_pos.grad).all())
