"""The two fused renderer kernels (csrc/gbuffer_fused.cu) against the step-by-step composition they replace -- the reference's
render_layer / shade / render_mesh flow (render/render.py:236-285, 55-63, 100-118, 144-186, 352-433) restated here with the
repo's single-attribute interpolate operator and plain torch -- forward buffers and gradients."""
import numpy as np
import pytest
import torch

from _device import DEVICE               # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu
D = DEVICE


def _scene(B=2, H=48, W=40, seed=0):
    from gshell_b200 import synthetic
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.grids import bcc_tet_grid
    from gshell_b200.render import mesh, raster, renderutils as ru
    v, t = bcc_tet_grid(6)
    g = torch.Generator().manual_seed(seed)
    pos = ((torch.tensor(v) - 0.5) * 2.0).to(D)
    sdf = (pos.norm(dim=1) - 0.7 + 0.05 * (torch.rand(v.shape[0], generator=g).to(D) - 0.5))
    msdf = pos[:, 1] + 0.3
    va, fa, _, _, _, ex = GShell_Tets(index_dtype=torch.int32, with_tangents=False)(pos, sdf, msdf, torch.tensor(t).to(D))
    m = mesh.auto_normals(mesh.Mesh(va.detach(), fa))
    mvp, campos = synthetic.random_cameras(B, (H, W), D, np.random.RandomState(seed))
    clip = ru.xfm_points(m.v_pos[None], mvp)
    rast, db = raster.rasterize(clip, fa, (H, W))
    assert float((rast[..., 3] > 0).float().mean()) > 0.1
    return m, ex["msdf"].detach(), clip, rast.detach(), db.detach(), campos


def test_gbuffer_matches_separate_interpolations():
    from gshell_b200.render import raster, render, util
    m, msdf, clip, rast, db, _ = _scene()
    tri = m.t_pos_idx.int()

    def reference(v_pos, v_nrm, ms):
        gb_pos, _ = raster.interpolate(v_pos[None], rast, tri)
        v0, v1, v2 = (v_pos[m.t_pos_idx[:, k].long()] for k in range(3))
        fn = util.safe_normalize(torch.linalg.cross(v1 - v0, v2 - v0))
        fidx = torch.arange(fn.shape[0], dtype=torch.int32, device=D)[:, None].repeat(1, 3)
        gb_geo, _ = raster.interpolate(fn[None], rast, fidx)
        gb_n, _ = raster.interpolate(v_nrm[None], rast, tri)
        with torch.no_grad():
            eps = 0.00001
            cp, cd = raster.interpolate(clip, rast, tri, rast_db=db)
            z0 = torch.clamp(cp[..., 2:3], min=eps) / torch.clamp(cp[..., 3:4], min=eps)
            z1 = torch.clamp(cp[..., 2:3] + torch.abs(cd[..., 2:3]), min=eps) / torch.clamp(cp[..., 3:4] + torch.abs(cd[..., 3:4]), min=eps)
            depth = torch.cat((z0, torch.abs(z1 - z0)), -1)
        mi, _ = raster.interpolate(ms.reshape(-1)[None, :, None], rast, tri)
        return gb_pos, gb_n, gb_geo, depth, mi

    la = [t.clone().requires_grad_() for t in (m.v_pos, m.v_nrm, msdf)]
    lb = [t.clone().requires_grad_() for t in (m.v_pos, m.v_nrm, msdf)]
    want = reference(*la)
    got = render._GBuffer.apply(lb[0], lb[1], lb[2], rast, db, m.t_pos_idx, clip)
    for name, a, b in zip(("pos", "normal", "geometric_normal", "depth", "msdf"), got, want):
        assert a.shape == b.shape, name
        assert float((a - b).abs().max()) <= 1e-5 * max(1.0, float(b.abs().max())), (name, float((a - b).abs().max()))
    gen = torch.Generator().manual_seed(1)
    ws = [torch.randn(w.shape, generator=gen).to(D) for w in want]
    pa = sum((w * x).sum() for k, (w, x) in enumerate(zip(ws, want)) if k != 3)
    pb = sum((w * x).sum() for k, (w, x) in enumerate(zip(ws, got)) if k != 3)
    ga = torch.autograd.grad(pa, la)
    gb = torch.autograd.grad(pb, lb)
    for name, a, b in zip(("v_pos", "v_nrm", "msdf"), gb, ga):
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()), (name, float((a - b).abs().max()), float(b.abs().max()))


@pytest.mark.parametrize("bsdf,composite", [("pbr", True), ("pbr", False), ("diffuse", True), ("white", True), ("override", True)])
def test_compose_matches_torch_composition(bsdf, composite):
    from gshell_b200.render import render, util
    m, msdf, clip, rast, db, _ = _scene(seed=3)
    B, H, W, _ = rast.shape
    g = torch.Generator().manual_seed(5)
    R = lambda *s: torch.rand(*s, generator=g).to(D)        # noqa: E731
    gb_nrm = torch.nn.functional.normalize(R(B, H, W, 3) - 0.5, dim=-1)
    tex, tex_j = R(B, H, W, 6), R(B, H, W, 6)
    sh_nrm, geo = R(B, H, W, 3), R(B, H, W, 3)
    depth = R(B, H, W, 2)
    diff, spec, col, mimg = R(B, H, W, 3), R(B, H, W, 3), R(B, H, W, 3), R(B, H, W, 1) - 0.5
    bg = R(B, H, W, 3)
    jitter = (torch.randn(B, H, W, 2, generator=g) * 0.02).to(D)          # larger than the renderer's 0.005: taps really move
    mode = {"pbr": 0, "diffuse": 1, "override": 2, "white": 3}[bsdf]

    def reference(gn, tx, txj, df, sp, cl, mi):
        mask = (rast[..., -1:] > 0).float()
        uv = (util.pixel_grid(W, H, device=D)[None] + jitter).contiguous()
        mask_tap = util.bilinear_tap(mask, uv)
        kd, ks = tx[..., 0:3], tx[..., 3:6]
        kd_grad = torch.abs(txj[..., 0:3] - kd)
        ks_grad = torch.abs(txj[..., 3:6] - ks) * torch.tensor([0.0, 1.0, 1.0], device=D)
        nrm_grad = torch.abs(util.bilinear_tap(gn, uv) - gn) * (mask * mask_tap)
        if bsdf == "pbr":
            kd_b = kd * (1.0 - ks[..., 2:3]); shaded = df * kd_b + sp
        elif bsdf == "diffuse":
            kd_b = kd; shaded = df * kd
        elif bsdf == "white":
            kd_b = torch.ones_like(kd); shaded = df * kd_b
        else:
            kd_b = kd; shaded = cl
        one = torch.ones_like(kd[..., 0:1])
        bufs = {"shaded": shaded, "z_grad": torch.cat((depth, torch.zeros_like(one)), -1), "normal": sh_nrm, "geometric_normal": geo,
                "kd": kd_b, "ks": ks, "kd_grad": kd_grad, "ks_grad": ks_grad, "normal_grad": nrm_grad}
        if bsdf != "override":
            bufs["diffuse_light"], bufs["specular_light"] = df, sp
        out = {}
        for k, x in bufs.items():
            x4 = torch.cat((x, one), -1)
            if composite:
                back = torch.cat((bg, torch.zeros_like(one)), -1) if k == "shaded" else torch.zeros_like(x4)
                x4 = torch.lerp(back, x4, mask)
            out[k] = x4
        # one-channel buffer: the reference's composite_buffer uses its last channel -- the value itself -- as alpha
        out["msdf_image"] = torch.lerp(torch.zeros_like(mi), torch.ones_like(mi), mask * mi) if composite else mi
        return out

    names = ("gb_nrm", "tex", "tex_j", "diff", "spec", "col", "msdf_img")
    la = [t.clone().requires_grad_() for t in (gb_nrm, tex, tex_j, diff, spec, col, mimg)]
    lb = [t.clone().requires_grad_() for t in (gb_nrm, tex, tex_j, diff, spec, col, mimg)]
    want = reference(*la)
    outs = render._Compose.apply(rast, jitter, lb[0], lb[1], lb[2], sh_nrm, geo, depth, None if bsdf == "override" else lb[3],
                                 None if bsdf == "override" else lb[4], lb[5] if bsdf == "override" else None, lb[6], bg, mode, composite)
    got = {k: o for k, o in zip(render._OUT_KEYS, outs) if o is not None}
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape, k
        # normal_grad goes through a bilinear tap: grid_sample and the kernel round the four weights differently (1e-5 level)
        tol = 2e-5 if k == "normal_grad" else 2e-6
        assert float((got[k] - want[k]).abs().max()) <= tol * max(1.0, float(want[k].abs().max())), (k, float((got[k] - want[k]).abs().max()))
    gen = torch.Generator().manual_seed(2)
    keys = [k for k in render._GRAD_KEYS if k in want]
    ws = {k: torch.randn(want[k].shape, generator=gen).to(D) for k in keys}
    pa = sum((ws[k] * want[k]).sum() for k in keys)
    pb = sum((ws[k] * got[k]).sum() for k in keys)
    used = [i for i, n in enumerate(names) if not (bsdf == "override" and n in ("diff", "spec")) and not (bsdf != "override" and n == "col")
            and not (bsdf == "white" and n == "spec")]
    ga = torch.autograd.grad(pa, [la[i] for i in used], allow_unused=True)
    gb = torch.autograd.grad(pb, [lb[i] for i in used], allow_unused=True)
    for i, a, b in zip(used, gb, ga):
        if b is None:
            assert a is None or float(a.abs().max()) == 0, names[i]
            continue
        assert a is not None, names[i]
        assert float((a - b).abs().max()) <= 5e-5 * max(1e-6, float(b.abs().max())), (names[i], float((a - b).abs().max()), float(b.abs().max()))


def test_render_mesh_buffers_and_keys():
    """API surface of render_mesh / render_layer: the reference's buffer names and channel counts."""
    from gshell_b200 import synthetic
    from gshell_b200.geometry.gshell_tets_geometry import default_flags
    from gshell_b200.render import light, optixutils as ou, render
    m, msdf, clip, rast, db, campos = _scene(seed=7)
    B, H, W, _ = rast.shape
    mvp, campos = synthetic.random_cameras(B, (H, W), D, np.random.RandomState(7))
    m.material = {"kd_ks": synthetic.LeafMaterialField(B, H, W, D), "bsdf": "pbr"}
    lgt = light.create_trainable_env_rnd(16, device=D)
    bufs = render.render_mesh(default_flags(n_samples=2), None, m, mvp, campos, lgt, [H, W], background=torch.rand(B, H, W, 3, device=D),
                              optix_ctx=ou.OptiXContext(), shadow_scale=0.0, use_uv=False, extra_dict={"msdf": msdf})
    assert {"shaded", "z_grad", "normal", "geometric_normal", "kd", "ks", "kd_grad", "ks_grad", "normal_grad", "diffuse_light",
            "specular_light", "msdf_image", "visible_triangles"} == set(bufs)
    for k, v in bufs.items():
        if k == "visible_triangles":
            assert v.dtype == torch.int64 and v.dim() == 1
        else:
            assert v.shape == (B, H, W, 1 if k == "msdf_image" else 4), k
            assert torch.isfinite(v).all(), k
    cov = (rast[..., 3] > 0)
    # alpha = coverage, softened across silhouettes by the antialiasing pass: 1 well inside, fractional on the outline
    a = bufs["shaded"][..., 3]
    assert bool((a[cov] > 0).all()) and float(a.max()) <= 1.0 and float(a.min()) >= 0.0
    assert int(((a > 0) & ~cov).sum()) > 0 and int(((a > 0) & ~cov).sum()) < 0.2 * int(cov.sum())
    render.antialias_enabled = False
    try:
        hard = render.render_mesh(default_flags(n_samples=2), None, m, mvp, campos, lgt, [H, W], optix_ctx=ou.OptiXContext(),
                                  shadow_scale=0.0, use_uv=False, extra_dict={"msdf": msdf})
    finally:
        render.antialias_enabled = True
    assert torch.equal(hard["shaded"][..., 3] > 0, cov) and torch.equal(hard["kd"][..., 3] > 0, cov)
