"""End-to-end: extraction -> normals -> clip transform -> rasterise -> interpolate -> shading normal -> MC shading
-> composite -> image loss, CUDA product path vs the composition of the oracles, forward image and gradients
w.r.t. SDF / mSDF / vertex positions / material / light."""
import numpy as np
import pytest
import torch

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("heavy", [False, True], ids=["plain", "shadows+denoiser"])
def test_train_step_matches_oracle_composition(heavy):
    """heavy: the two stages that dominate the benchmark step are switched on -- every shadow ray traced (shadow_scale = 1,
    oracle: brute-force ray/triangle tests against the extracted mesh) and the cross-bilateral denoiser on diffuse and
    specular light (oracle: the restated filter; its depth guide, a non-differentiable rasteriser output, is taken from the
    product's own z-buffer)."""
    from gshell_b200 import synthetic
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.grids import bcc_tet_grid
    from gshell_b200.render import light, mesh, render
    from gshell_b200.render import renderutils as ru
    from gshell_b200.geometry.gshell_tets_geometry import default_flags
    from oracle import mt_oracle, raster_oracle, shade_oracle as so
    d = device()
    B, H, W, n = 2, 40, 40, 2
    v, t = bcc_tet_grid(6)
    g = torch.Generator().manual_seed(0)
    pos = (torch.tensor(v) - 0.5) * 2.0
    sdf = pos.norm(dim=1) - 0.7 + 0.05 * (torch.rand(v.shape[0], generator=g) - 0.5)
    msdf = pos[:, 1] + 0.3 + 0.05 * torch.rand(v.shape[0], generator=g)
    tets = torch.tensor(t)
    rng = np.random.RandomState(0)
    mvp, campos = synthetic.random_cameras(B, (H, W), "cpu", rng)
    img, bg = synthetic.random_target(B, (H, W), "cpu", g)
    tex = torch.cat([torch.rand(B, H, W, 3, generator=g),
                     torch.stack([torch.zeros(B, H, W), 0.4 + 0.5 * torch.rand(B, H, W, generator=g), torch.rand(B, H, W, generator=g)], -1)], -1)
    base = torch.rand(16, 32, 3, generator=g) * 0.5 + 0.25
    perms = torch.argsort(torch.rand(32768, n * n, generator=g), dim=-1).int()

    # ------------------------------- product path (CUDA) ---------------------------------------------------
    gl = [x.clone().to(d).requires_grad_() for x in (pos, sdf, msdf, tex, base)]
    gva, gfa, _, _, _, gex = GShell_Tets(index_dtype=torch.int32, with_tangents=False)(gl[0], gl[1], gl[2], tets.to(d))
    m = mesh.auto_normals(mesh.Mesh(gva, gfa, material={"kd_ks": type("F", (), {"sample": lambda self, p: gl[3]})(), "bsdf": "pbr"}))
    lgt = light.EnvironmentLight(gl[4])
    FLAGS = default_flags(n_samples=n)
    render.rnd_seed = 0
    from gshell_b200.render.optixutils import ops as ouops
    ouops._EnvShade._random_perm[(n, str(d))] = perms.to(d).contiguous()
    from gshell_b200.render import optixutils as ou
    octx = ou.OptiXContext()
    den = None
    render.antialias_enabled = False        # the oracle composition has no antialiasing stage (tests/test_antialias_gpu.py covers it)
    if heavy:
        from gshell_b200.denoiser.denoiser import BilateralDenoiser
        ou.optix_build_bvh(octx, gva, gfa, rebuild=1)
        den = BilateralDenoiser(0.5)
    bufs = render.render_mesh(FLAGS, None, m, mvp.to(d), campos.to(d), lgt, [H, W], spp=1, msaa=True, background=bg.to(d),
                              optix_ctx=octx, bsdf=None, denoiser=den, shadow_scale=1.0 if heavy else 0.0, use_uv=False,
                              extra_dict={"msdf": gex["msdf"]})
    render.antialias_enabled = True
    g_img = bufs["shaded"][..., 0:3]
    ti = img.to(d)
    g_loss = ru.image_loss(g_img * ti[..., 3:], ti[..., 0:3] * ti[..., 3:], loss="l1", tonemapper="log_srgb") + \
        0.5 * (bufs["msdf_image"][..., 0:1].clamp(min=0) * (ti[..., 3:] == 0).float()).abs().mean()
    g_loss.backward()
    zdz = bufs["z_grad"][..., 0:2].detach().cpu()

    # ------------------------------- oracle composition (CPU) ---------------------------------------------
    ol = [x.clone().requires_grad_() for x in (pos, sdf, msdf, tex, base)]
    va, fa, _, _, _, ex = mt_oracle.gshell_marching_tets(ol[0], ol[1], ol[2], tets, unique_mode="packed", with_tangents=False)
    assert torch.equal(gfa.cpu().long(), fa)
    vn = mt_oracle.smooth_normals(va, fa)
    clip = so.xfm_points(va[None], mvp)
    rast = raster_oracle.rasterize(clip, fa, H, W)
    gb_pos = raster_oracle.interpolate(va[None], rast, fa)
    p0, p1, p2 = va[fa[:, 0]], va[fa[:, 1]], va[fa[:, 2]]
    fn = torch.linalg.cross(p1 - p0, p2 - p0)
    fn = fn / torch.sqrt(torch.clamp((fn * fn).sum(-1, keepdim=True), min=1e-20))
    fidx = torch.arange(fa.shape[0])[:, None].repeat(1, 3)
    gb_gn = raster_oracle.interpolate(fn[None], rast, fidx)
    gb_n = raster_oracle.interpolate(vn[None], rast, fa)
    vp = campos[:, None, None, :]
    tng = torch.linalg.cross(torch.randn(B, H, W, 3, generator=g), gb_n.detach())
    sh_n = so.prepare_shading_normal(gb_pos, vp, None, gb_n, tng, gb_gn, True, True)
    pdf, rows, cols = so.light_pdf_tables(ol[4].detach())
    kd, ks = ol[3][..., 0:3], ol[3][..., 3:6]
    vis_fn = None
    if heavy:
        from test_shade_gpu import _brute_force_visibility
        vis_fn = _brute_force_visibility(va.detach(), fa)
    od, os_ = so.env_shade(rast[..., 3].detach(), gb_pos + sh_n * 0.001, gb_pos, sh_n, vp, kd, ks, ol[4], pdf, rows, cols, perms,
                           bsdf=0, n_samples_x=n, rnd_seed=0, shadow_scale=1.0 if heavy else 0.0, visibility=vis_fn)
    if heavy:
        guide = sh_n / torch.sqrt(torch.clamp((sh_n * sh_n).sum(-1, keepdim=True), min=1e-20))
        od = so.bilateral_denoiser(od, guide, zdz, den.sigma)
        os_ = so.bilateral_denoiser(os_, guide, zdz, den.sigma)
    shaded = od * kd * (1.0 - ks[..., 2:3]) + os_
    cov = (rast[..., 3:4] > 0).float().detach()
    o_img = torch.lerp(bg, shaded, cov)
    o_msdf = raster_oracle.interpolate(ex["msdf"][None, :, None], rast, fa)
    o_loss = so.image_loss(o_img * img[..., 3:], img[..., 0:3] * img[..., 3:], "l1", "log_srgb") + \
        0.5 * (o_msdf.clamp(min=0) * (img[..., 3:] == 0).float()).abs().mean()
    o_loss.backward()

    # forward image: bulk within 1e-4 relative; pixels where a discrete sample decision flipped are bounded
    a, b = g_img.detach().cpu(), o_img.detach()
    rel = (a - b).abs() / b.abs().clamp(min=1e-3)
    assert float(rel.median()) < 1e-5 and float((rel > 1e-4).float().mean()) < 0.02, (float(rel.median()), float((rel > 1e-4).float().mean()))
    assert abs(float(g_loss) - float(o_loss)) <= 1e-4 * abs(float(o_loss))
    for name, x, y in zip(("pos", "sdf", "msdf", "material", "light"), gl, ol):
        l2 = float((x.grad.cpu() - y.grad).norm() / y.grad.norm().clamp(min=1e-20))
        print("pipeline grad", name, "rel L2", l2)
        assert l2 < 5e-3, (name, l2)


@pytest.mark.parametrize("kind", ["tets", "flexicubes", "flexicubes_sdf_mlp", "tets_mlp_material"])
def test_geometry_tick_runs_and_optimises(kind, tmp_path):
    """API-level smoke of the training surface: GShell*Geometry.tick() -> backward -> Adam for a few iterations; loss finite,
    every parameter group receives a finite gradient, dict keys of getMesh() as the reference's."""
    from gshell_b200 import synthetic
    from gshell_b200.denoiser.denoiser import BilateralDenoiser
    from gshell_b200.geometry.gshell_tets_geometry import GShellTetsGeometry, default_flags
    from gshell_b200.geometry.gshell_flexicubes_geometry import GShellFlexiCubesGeometry
    from gshell_b200.grids import save_tets_npz
    from gshell_b200.render import light
    from gshell_b200.render import renderutils as ru
    d = device()
    torch.manual_seed(0)
    FLAGS = default_flags(n_samples=2, sphere_init=True, use_sdf_mlp=kind.endswith("_sdf_mlp"), use_msdf_mlp=kind.endswith("_msdf_mlp"), sdf_mlp_pretrain_steps=400, d_hidden=64,
                          n_hidden=2, skip_in=[1])
    if kind.startswith("tets"):
        npz = str(tmp_path / "tets.npz")
        save_tets_npz(npz, 10)
        geo = GShellTetsGeometry(64, 2.0, FLAGS, tet_init_file=npz, device=d)
        # use_msdf_mlp (reference gshell_tets_geometry.py:118-136, 199-202): the mSDF comes from a field MLP, `msdf` is a placeholder
        params = [geo.sdf] + (list(geo.msdf_net.parameters()) if FLAGS.use_msdf_mlp else [geo.msdf]) + [geo.deform]
    else:
        geo = GShellFlexiCubesGeometry(12, 2.0, FLAGS, device=d)
        params = (list(geo.sdf_net.parameters()) if FLAGS.use_sdf_mlp else [geo.sdf]) + [geo.msdf, geo.deform, geo.per_cube_weights]
    B, res = 2, [48, 48]
    rng = np.random.RandomState(1)
    if kind == "tets_mlp_material":
        # the reference's learned material: hash-grid encoding + MLP sampled at the G-buffer positions (render/mlptexture.py)
        from gshell_b200.render.mlptexture import MLPTexture3D
        aabb = torch.tensor([[-1.2] * 3, [1.2] * 3], device=d)
        mn = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.08, 0.0], device=d)
        mx = torch.tensor([1.0, 1.0, 1.0, 0.0, 1.0, 1.0], device=d)
        mat = MLPTexture3D(aabb, channels=6, min_max=[mn, mx])
        mat_params = list(mat.parameters())
    else:
        mat = synthetic.LeafMaterialField(B, res[0], res[1], d)
        mat_params = [mat.tex]
    material = {"kd_ks": mat, "bsdf": "pbr"}
    lgt = light.create_trainable_env_rnd(16, device=d)
    mvp, campos = synthetic.random_cameras(B, res, d, rng)
    img, bg = synthetic.random_target(B, res, d)
    target = {"mvp": mvp, "campos": campos, "img": img, "background": bg, "resolution": res, "spp": 1}
    mesh_dict = geo.getMesh(material)
    assert {"imesh", "sdf", "msdf", "msdf_watertight", "msdf_boundary", "n_verts_watertight"} <= set(mesh_dict)
    assert mesh_dict["imesh"].v_pos.shape[0] > 0 and mesh_dict["imesh"].v_nrm.shape == mesh_dict["imesh"].v_pos.shape
    opt = torch.optim.Adam(params + mat_params + [lgt.base], lr=1e-3)
    den = BilateralDenoiser().to(d)
    loss_fn = lambda a, b: ru.image_loss(a, b, loss="l1", tonemapper="log_srgb")    # noqa: E731
    for it in (0, 500, 1500):
        lgt.update_pdf()
        opt.zero_grad()
        il, dl, rl = geo.tick(None, target, lgt, material, loss_fn, it, den)
        total = il + dl + rl
        assert torch.isfinite(total)
        total.backward()
        for p in params + mat_params + [lgt.base]:
            assert p.grad is not None and torch.isfinite(p.grad).all()
        assert all(float(p.grad.abs().sum()) > 0 for p in mat_params)
        assert float(params[0].grad.abs().sum()) > 0 and float(lgt.base.grad.abs().sum()) > 0
        opt.step()


@pytest.mark.parametrize("cfg", ["polycam_mc_128", "deepfashion_mc_80"])
def test_baseline_config_shapes_run(cfg, tmp_path):
    """BASELINE.json configs[1] and [2] (parity-test cases, not bench lines): one full training iteration at the configured
    shapes -- '128' tet grid (BCC N=52) resp. 80^3 G-FlexiCubes grid, 4 views @ 512^2, n_samples = 8 -- finite loss and
    gradients, mesh non-empty, index ranges valid."""
    from gshell_b200 import synthetic
    from gshell_b200.denoiser.denoiser import BilateralDenoiser
    from gshell_b200.geometry.gshell_tets_geometry import GShellTetsGeometry, default_flags
    from gshell_b200.geometry.gshell_flexicubes_geometry import GShellFlexiCubesGeometry
    from gshell_b200.grids import save_tets_npz
    from gshell_b200.render import light
    from gshell_b200.render import renderutils as ru
    d = device()
    torch.manual_seed(0)
    FLAGS = default_flags(n_samples=8, sphere_init=True)
    if cfg == "polycam_mc_128":
        npz = str(tmp_path / "tets.npz")
        save_tets_npz(npz, 52)
        geo = GShellTetsGeometry(128, 2.0, FLAGS, tet_init_file=npz, device=d)
    else:
        geo = GShellFlexiCubesGeometry(80, 2.0, FLAGS, device=d)
    B, res = 4, [512, 512]
    mat = synthetic.LeafMaterialField(B, res[0], res[1], d)
    lgt = light.create_trainable_env_rnd(256, device=d)
    mvp, campos = synthetic.random_cameras(B, res, d, np.random.RandomState(2))
    img, bg = synthetic.random_target(B, res, d)
    target = {"mvp": mvp, "campos": campos, "img": img, "background": bg, "resolution": res, "spp": 1}
    lgt.update_pdf()
    il, dl, rl = geo.tick(None, target, lgt, {"kd_ks": mat, "bsdf": "pbr"}, lambda a, b: ru.image_loss(a, b, loss="l1", tonemapper="log_srgb"),
                          1200, BilateralDenoiser())
    total = il + dl + rl
    assert torch.isfinite(total)
    total.backward()
    for p in (geo.sdf, geo.msdf, geo.deform, mat.tex, lgt.base):
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert float(geo.sdf.grad.abs().sum()) > 0 and float(mat.tex.grad.abs().sum()) > 0
