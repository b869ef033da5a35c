"""GPU parity of the G-buffer / shading operators (through the C ABI) against the reference-pinned
goldens (tests/golden/shade_*.npz) and the oracle (oracle/shade_oracle.py).
Tolerance: 1e-4 relative (north_star) unless a test states and justifies otherwise."""
import os

import numpy as np
import pytest
import torch

from _device import DEVICE, device, synchronize      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def dev():
    return device()


def load(name):
    z = np.load(os.path.join(G, f"shade_{name}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_close(got, want, tol=1e-4, what=""):
    got, want = got.detach().cpu().float(), want.detach().cpu().float()
    scale = want.abs().max().clamp(min=1e-6)
    err = (got - want).abs().max()
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def test_xfm_points_golden():
    import gshell_b200.render.renderutils as ru
    g = load("xfm")
    pts = g["points"].clone().to(dev()).requires_grad_()
    out = ru.xfm_points(pts, g["matrix"].to(dev()))
    rel_close(out, g["out"], 1e-5, "xfm out")
    (out * g["w"].to(dev())).sum().backward()
    rel_close(pts.grad, g["g_points"], 1e-5, "xfm grad")
    # batched points
    pb = g["points"].expand(3, -1, -1).contiguous().to(dev()).requires_grad_()
    outb = ru.xfm_points(pb, g["matrix"].to(dev()))
    rel_close(outb, g["out"], 1e-5)
    (outb * g["w"].to(dev())).sum().backward()
    rel_close(pb.grad.sum(0, keepdim=True), g["g_points"], 1e-5)


@pytest.mark.parametrize("name", ["nrm_plain", "nrm_perturbed"])
def test_prepare_shading_normal_golden(name):
    import gshell_b200.render.renderutils as ru
    g = load(name)
    keys = ["pos", "view_pos", "smooth_nrm", "smooth_tng", "geom_nrm"] + (["perturbed_nrm"] if "perturbed_nrm" in g else [])
    L = {k: g[k].clone().to(dev()).requires_grad_() for k in keys}
    out = ru.prepare_shading_normal(L["pos"], L["view_pos"], L.get("perturbed_nrm"), L["smooth_nrm"], L["smooth_tng"],
                                    L["geom_nrm"], two_sided_shading=True, opengl=True)
    rel_close(out, g["out"], 1e-4, "normal out")
    (out * g["w"].to(dev())).sum().backward()
    for k in keys:
        rel_close(L[k].grad, g["g_" + k], 1e-4, "grad " + k)


def test_image_loss_golden():
    import gshell_b200.render.renderutils as ru
    g = load("loss")
    for loss, tm in (("l1", "none"), ("l1", "log_srgb"), ("mse", "log_srgb"), ("smape", "none"), ("relmse", "none"), ("mse", "none")):
        img, tgt = g["img"].clone().to(dev()).requires_grad_(), g["target"].clone().to(dev()).requires_grad_()
        val = ru.image_loss(img, tgt, loss=loss, tonemapper=tm)
        rel_close(val, g[f"{loss}_{tm}"], 1e-5, f"{loss}/{tm}")
        val.backward()
        rel_close(img.grad, g[f"{loss}_{tm}_g_img"], 1e-4, f"{loss}/{tm} g_img")
        rel_close(tgt.grad, g[f"{loss}_{tm}_g_target"], 1e-4, f"{loss}/{tm} g_target")


@pytest.mark.parametrize("name", ["denoise_s06", "denoise_s10"])
def test_bilateral_denoiser_golden(name):
    import gshell_b200.render.optixutils as ou
    g = load(name)
    col = g["col"].clone().to(dev()).requires_grad_()
    nrm, zdz, sigma = g["nrm"].to(dev()), g["zdz"].to(dev()), float(g["sigma"])
    out = ou.bilateral_denoiser(col, nrm, zdz, sigma)
    rel_close(out, g["out"], 1e-4, "denoise out")
    (out * g["w"].to(dev())).sum().backward()
    rel_close(col.grad, g["g_col"], 1e-4, "denoise grad")
    # strided channel-slice views of one [B,H,W,8] tensor (how render.py feeds it) + fused pair
    packed = torch.cat([g["col"], g["nrm"], g["zdz"]], -1).to(dev())
    out2 = ou.bilateral_denoiser(packed[..., 0:3], packed[..., 3:6], packed[..., 6:8], sigma)
    assert torch.equal(out2, out.detach())
    ca = g["col"].clone().to(dev()).requires_grad_()
    cb = (g["col"].flip(-1) * 0.5).to(dev()).requires_grad_()
    oa, ob = ou.bilateral_denoiser_pair(ca, cb, nrm, zdz, sigma)
    assert torch.allclose(oa, out.detach(), rtol=1e-6, atol=1e-7)
    ob_single = ou.bilateral_denoiser(cb.detach(), nrm, zdz, sigma)
    assert torch.allclose(ob, ob_single, rtol=1e-6, atol=1e-7)
    (oa * g["w"].to(dev())).sum().backward()
    rel_close(ca.grad, g["g_col"], 1e-4, "pair grad")


def _shade_inputs(B, H, W, seed, lh=16, lw=32, rough_min=0.08):
    g = torch.Generator().manual_seed(seed)
    R = lambda *s: torch.rand(*s, generator=g)        # noqa: E731
    N = lambda *s: torch.randn(*s, generator=g)       # noqa: E731
    nrm = torch.nn.functional.normalize(N(B, H, W, 3), dim=-1)
    pos = N(B, H, W, 3) * 0.3
    view = (pos.mean((1, 2), keepdim=True) + 3.0 * torch.nn.functional.normalize(N(B, 1, 1, 3), dim=-1))
    # make most normals face the camera, as a rasterised G-buffer would
    facing = ((view - pos) * nrm).sum(-1, keepdim=True) > 0
    nrm = torch.where(facing, nrm, -nrm)
    kd = R(B, H, W, 3)
    ks = torch.stack([torch.zeros(B, H, W), rough_min + (0.98 - rough_min) * R(B, H, W), R(B, H, W)], -1)
    mask = (R(B, H, W) > 0.2).float()
    light = R(lh, lw, 3) * 0.5 + 0.25
    light[2:4, 5:9] = 15.0
    return mask, pos, nrm, view, kd, ks, light


@pytest.mark.parametrize("bsdf,n,seed,rough_min", [("pbr", 4, 1, 0.3), ("pbr", 3, 2, 0.3), ("diffuse", 4, 3, 0.08),
                                                    ("pbr", 4, 4, 0.08)])
def test_env_shade_vs_oracle(bsdf, n, seed, rough_min):
    """Parity against oracle/shade_oracle.py::env_shade, which is itself pinned to the reference's own kernel.cu compiled for
    the CPU (tests/test_oracle_env_shade_ref.py).
    Per-sample discrete decisions (nearest light texel, CDF bin, lobe choice) can flip between CPU libm and
    CUDA libm on a 1-ulp difference, moving one of the 2n^2 samples of a pixel: such pixels are reported and
    bounded (<1%), every other covered pixel must agree to 1e-4 relative.
    Conditioning: the GGX lobe D = a2 / (pi ((c a2 - c) c + 1)^2) amplifies the fp32 rounding of c = n.h by
    ~2/(a2 + 1 - c^2); at the reference's minimum roughness 0.08 (a2 = 4e-5) a 1-ulp difference in c moves a
    highlight pixel by ~1e-3 relative in ANY fp32 implementation (FMA contraction alone does it).  The strict
    1e-4 bar is therefore asserted for roughness >= 0.3; the full-range case asserts the bulk (median) and
    bounds the tail.  The tail bounds are MEASURED, not argued: tests/test_oracle_env_shade_conditioning.py runs the
    reference's own kernel.cu, compiled with and without contraction / fast math, on exactly these inputs -- against
    itself it moves 3.5 - 6.7 % of the pixels by more than 1e-4 (largest 0.7e-3 - 1.9e-3) and its gradients by up to
    1.2e-3 relative L2.  Bounds here: 8 % of the pixels, largest 5e-3, gradients 1e-2 (B200: 4.9 %, 2.2e-3, 2.1e-3)."""
    import gshell_b200.render.optixutils as ou
    from oracle import shade_oracle as so
    B, H, W = 2, 24, 20
    mask, pos, nrm, view, kd, ks, light = _shade_inputs(B, H, W, seed, rough_min=rough_min)
    strict = rough_min >= 0.3 or bsdf != "pbr"
    pdf, rows, cols = so.light_pdf_tables(light)
    perms = torch.argsort(torch.rand(32768, n * n, generator=torch.Generator().manual_seed(seed)), dim=-1).int()
    ib = ["pbr", "diffuse", "white"].index(bsdf)
    ol = [t.clone().requires_grad_() for t in (pos, nrm, kd, ks, light)]
    od, os_ = so.env_shade(mask, ol[0] + 0.001 * ol[1], ol[0], ol[1], view, ol[2], ol[3], ol[4], pdf, rows, cols, perms,
                           bsdf=ib, n_samples_x=n, rnd_seed=17 + seed, shadow_scale=0.0)
    d = dev()
    gl = [t.clone().to(d).requires_grad_() for t in (pos, nrm, kd, ks, light)]
    gd, gs = ou.optix_env_shade(ou.OptiXContext(), mask.to(d), (gl[0] + 0.001 * gl[1]).detach(), gl[0], gl[1], view.to(d),
                                gl[2], gl[3], gl[4], pdf.to(d), rows.to(d), cols.to(d), BSDF=bsdf, n_samples_x=n,
                                rnd_seed=17 + seed, shadow_scale=0.0, perms=perms.to(d))
    cov = mask > 0
    assert float(gd.cpu()[~cov].abs().max()) == 0 and float(gs.cpu()[~cov].abs().max()) == 0
    stats = {}
    for name, got, want in (("diff", gd, od), ("spec", gs, os_)):
        got, want = got.detach().cpu(), want.detach()
        floor = 1e-3 * want[cov].abs().mean().clamp(min=1e-8)
        rel = ((got - want).abs() / want.abs().clamp(min=floor))[cov]
        bad = (rel > 1e-4).float().mean().item()
        stats[name] = (rel.median().item(), bad, rel.max().item())
        assert rel.median() < 1e-5, (name, stats)
        if strict:
            assert bad < 0.01, (name, stats)
        else:
            assert bad < 0.08 and rel.max() < 5e-3, (name, stats)
    print("env_shade parity", bsdf, n, rough_min, stats)
    gen = torch.Generator().manual_seed(99)
    wd, ws = torch.randn(od.shape, generator=gen), torch.randn(os_.shape, generator=gen)
    (od * wd).sum().add((os_ * ws).sum()).backward()
    ((gd * wd.to(d)).sum() + (gs * ws.to(d)).sum()).backward()
    for name, a, b in zip(("pos", "nrm", "kd", "ks", "light"), gl, ol):
        if b.grad is None:
            assert a.grad is None or float(a.grad.abs().max()) == 0
            continue
        want, got = b.grad, a.grad.cpu()
        # gradients: same flip caveat; compare in aggregate (relative L2) and per element on the bulk
        l2 = (got - want).norm() / want.norm().clamp(min=1e-12)
        print("  grad", name, "rel L2", float(l2))
        assert l2 < (2e-3 if strict else 1e-2), (name, float(l2))
        floor = 1e-3 * want.abs().mean().clamp(min=1e-12)
        rel = (got - want).abs() / want.abs().clamp(min=floor)
        sel = want.abs() > floor
        if sel.any():
            assert rel[sel].median() < 1e-4, (name, float(rel[sel].median()))


@pytest.mark.parametrize("n,shadows", [(8, True), (16, False), (16, True)])
def test_env_shade_vs_compiled_reference_at_benchmark_sample_counts(n, shadows):
    """BASELINE.json's sample counts (n = 8: configs[1]/[2], n = 16: configs[3]) on a 64^2 crop, compared DIRECTLY with the
    reference's own integrator (envsampling/kernel.cu compiled unmodified for the CPU, oracle/_ref; shadow rays answered by its
    brute-force any-hit loop) -- forward and all five gradients, with and without occluders."""
    import gshell_b200.render.optixutils as ou
    from oracle import ref_env_shade as ref, shade_oracle as so
    B, H, W = 1, 64, 64
    mask, pos, nrm, view, kd, ks, light = _shade_inputs(B, H, W, 30 + n, rough_min=0.3, lh=32, lw=64)
    g = torch.Generator().manual_seed(n)
    pdf, rows, cols = so.light_pdf_tables(light)
    perms = torch.argsort(torch.rand(32768, n * n, generator=g), dim=-1).int()
    d = dev()
    verts = tris = None
    ctx = ou.OptiXContext()
    if shadows:
        # a soup of small triangles around the shaded points casts plenty of shadows
        c = torch.randn(300, 1, 3, generator=g) * 0.5
        verts = (c + 0.12 * torch.randn(300, 3, 3, generator=g)).reshape(-1, 3)
        tris = torch.arange(900, dtype=torch.int32).reshape(-1, 3)
        ou.optix_build_bvh(ctx, verts.to(d), tris.to(d), rebuild=1)
    ss = 1.0 if shadows else 0.0
    ro = pos + 0.001 * nrm
    a = (mask, ro, pos, nrm, view, kd, ks, light, pdf, rows, cols, perms)
    rd, rs = ref.env_shade_fwd(*a, bsdf=0, n_samples_x=n, rnd_seed=9, shadow_scale=ss, verts=verts, tris=tris)
    gen = torch.Generator().manual_seed(99)
    wd, ws = torch.randn(rd.shape, generator=gen), torch.randn(rs.shape, generator=gen)
    rgrads = ref.env_shade_bwd(*a, wd, ws, bsdf=0, n_samples_x=n, rnd_seed=9, shadow_scale=ss, verts=verts, tris=tris)
    gl = [t.clone().to(d).requires_grad_() for t in (pos, nrm, kd, ks, light)]
    gd, gs = ou.optix_env_shade(ctx, mask.to(d), ro.to(d), gl[0], gl[1], view.to(d), gl[2], gl[3], gl[4], pdf.to(d), rows.to(d),
                                cols.to(d), BSDF="pbr", n_samples_x=n, rnd_seed=9, shadow_scale=ss, perms=perms.to(d))
    cov = mask > 0
    if shadows:
        unsh, _ = ref.env_shade_fwd(*a, bsdf=0, n_samples_x=n, rnd_seed=9, shadow_scale=0.0)
        assert float(((unsh - rd).abs().sum(-1) > 1e-6)[cov].float().mean()) > 0.3, "scene casts too few shadows"
    for name, got, want in (("diff", gd, rd), ("spec", gs, rs)):
        got = got.detach().cpu()
        floor = 1e-3 * want[cov].abs().mean().clamp(min=1e-8)
        rel = ((got - want).abs() / want.abs().clamp(min=floor))[cov]
        assert rel.median() < 1e-5 and (rel > 1e-4).float().mean() < 0.01, (name, float(rel.median()), float((rel > 1e-4).float().mean()))
    ((gd * wd.to(d)).sum() + (gs * ws.to(d)).sum()).backward()
    for name, x, want in zip(("pos", "nrm", "kd", "ks", "light"), gl, rgrads):
        l2 = float((x.grad.cpu() - want).norm() / want.norm().clamp(min=1e-12))
        print("grad vs compiled reference", n, shadows, name, "rel L2", l2)
        # d_light sums ~10^6 per-sample terms per texel with fp32 atomics in launch order (the reference's CUDA build does the
        # same, kernel.cu:455; its CPU build here sums in pixel order): the sum's rounding alone is ~2e-3 of its norm
        assert l2 < (5e-3 if name == "light" else 2e-3), (name, l2)


def test_env_shade_understated_pixel_count_is_an_error(monkeypatch):
    """A ray list sized for fewer pixels than the mask holds loses rays; the library must say so (C ABI: every later traced
    call fails until gsb_env_shade_dropped_rays(1) acknowledges it) instead of returning a plausible image."""
    import gshell_b200.render.optixutils as ou
    from gshell_b200 import _lib
    from gshell_b200.render.optixutils import ops
    from oracle import shade_oracle as so
    d = dev()
    B, H, W, n = 1, 32, 32, 4
    mask, pos, nrm, view, kd, ks, light = _shade_inputs(B, H, W, 3, rough_min=0.3)
    mask = torch.ones_like(mask)
    pdf, rows, cols = so.light_pdf_tables(light)
    verts = torch.tensor([[-5.0, -5.0, 1.0], [5.0, -5.0, 1.0], [0.0, 5.0, 1.0]], device=d)
    ctx = ou.OptiXContext()
    ou.optix_build_bvh(ctx, verts, torch.tensor([[0, 1, 2]], dtype=torch.int32, device=d), rebuild=1)
    args = [t.to(d) for t in (mask, pos + 0.001 * nrm, pos, nrm, view, kd, ks, light, pdf, rows, cols)]
    _lib.lib.gsb_env_shade_dropped_rays(1)
    ou.optix_env_shade(ctx, *args, BSDF="pbr", n_samples_x=n, rnd_seed=1, shadow_scale=1.0)
    assert _lib.lib.gsb_env_shade_dropped_rays(0) == 0
    monkeypatch.setattr(ops, "_covered_pixels", lambda m: 16)          # 1024 pixels emit rays, the list holds 16 pixels' worth
    ops._scratch_cache.clear()
    ou.optix_env_shade(ctx, *args, BSDF="pbr", n_samples_x=n, rnd_seed=1, shadow_scale=1.0)
    synchronize()
    with pytest.raises(RuntimeError):
        ou.optix_env_shade(ctx, *args, BSDF="pbr", n_samples_x=n, rnd_seed=1, shadow_scale=1.0)
    assert _lib.lib.gsb_env_shade_dropped_rays(1) > 0
    monkeypatch.undo()
    ops._scratch_cache.clear()
    ou.optix_env_shade(ctx, *args, BSDF="pbr", n_samples_x=n, rnd_seed=1, shadow_scale=1.0)       # acknowledged: works again
    assert _lib.lib.gsb_env_shade_dropped_rays(0) == 0


def test_env_shade_white_furnace_full_res():
    """Size-independent property at the benchmark resolution: constant white probe + diffuse BSDF integrates the
    cosine lobe to 1 at every covered pixel (stratified MIS estimator, n=4 -> 32 samples)."""
    import gshell_b200.render.optixutils as ou
    from oracle import shade_oracle as so
    d = dev()
    B, H, W, n = 1, 1024, 1024, 4
    torch.manual_seed(0)
    nrm = torch.nn.functional.normalize(torch.randn(B, H, W, 3, device=d), dim=-1)
    nrm[..., 2] = nrm[..., 2].abs().clamp(min=0.05)      # camera-facing, as a rasterised + bent G-buffer normal is
    nrm = torch.nn.functional.normalize(nrm, dim=-1)      # (for N.V < 1e-6 the reference's bsdf_pdf returns 1, kernel.cu:382)
    pos = torch.zeros(B, H, W, 3, device=d)
    view = torch.tensor([[[[0.0, 0.0, 3.0]]]], device=d)
    light = torch.ones(64, 128, 3)
    pdf, rows, cols = so.light_pdf_tables(light)
    mask = torch.ones(B, H, W, device=d)
    kd = torch.rand(B, H, W, 3, device=d)
    ks = torch.rand(B, H, W, 3, device=d)
    diff, spec = ou.optix_env_shade(ou.OptiXContext(), mask, pos, pos, nrm, view, kd, ks, light.to(d), pdf.to(d),
                                    rows.to(d), cols.to(d), BSDF="diffuse", n_samples_x=n, rnd_seed=5, shadow_scale=0.0)
    assert torch.isfinite(diff).all() and float(spec.abs().max()) == 0
    assert abs(float(diff.mean()) - 1.0) < 0.01
    assert float((diff.mean(-1) - 1.0).abs().quantile(0.99)) < 0.5


def _brute_force_visibility(verts, tris):
    """Oracle stand-in for the OptiX shadow ray: Moeller-Trumbore against every triangle, t in (0, 1e16)."""
    v0 = verts[tris[:, 0]]; e1 = verts[tris[:, 1]] - v0; e2 = verts[tris[:, 2]] - v0

    def vis(o, d):
        o, d = o.detach(), d.detach()
        p = torch.linalg.cross(d[:, None, :].expand(-1, e2.shape[0], -1), e2[None].expand(d.shape[0], -1, -1))
        det = (e1[None] * p).sum(-1)
        ok = det != 0
        inv = 1.0 / torch.where(ok, det, torch.ones_like(det))
        t_ = o[:, None, :] - v0[None]
        u = (t_ * p).sum(-1) * inv
        q = torch.linalg.cross(t_, e1[None].expand_as(t_))
        v = (d[:, None, :] * q).sum(-1) * inv
        t = (e2[None] * q).sum(-1) * inv
        hit = ok & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 0) & (t < 1e16)
        return (~hit.any(1)).float()[:, None]
    return vis


def test_env_shade_shadow_rays_vs_oracle():
    """shadow_scale = 1: visibility from the uniform-grid occluder vs brute-force ray/triangle tests in the oracle.
    Geometry: an extracted G-Shell mesh; G-buffer points sit on that mesh (offset along the normal like render.py:131)."""
    import gshell_b200.render.optixutils as ou
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.grids import bcc_tet_grid
    from oracle import shade_oracle as so
    d = dev()
    v, t = bcc_tet_grid(7)
    g = torch.Generator().manual_seed(4)
    pos3 = (torch.tensor(v) - 0.5) * 2
    sdf = pos3.norm(dim=1) - 0.6 + 0.25 * (torch.rand(v.shape[0], generator=g) - 0.5)
    msdf = torch.rand(v.shape[0], generator=g) - 0.2
    va, fa, _, _, _, _ = GShell_Tets(index_dtype=torch.int32, with_tangents=False)(pos3.to(d), sdf.to(d), msdf.to(d), torch.tensor(t).to(d))
    va_c, fa_c = va.cpu(), fa.cpu().long()
    B, H, W, n = 1, 16, 16, 3
    sel = torch.randint(0, fa_c.shape[0], (B * H * W,), generator=g)
    bary = torch.rand(B * H * W, 3, generator=g); bary = bary / bary.sum(-1, keepdim=True)
    tri = va_c[fa_c[sel]]
    pos = (tri * bary[..., None]).sum(1).view(B, H, W, 3)
    fn = torch.nn.functional.normalize(torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), dim=-1).view(B, H, W, 3)
    view = torch.tensor([0.0, 0.0, 3.0]).view(1, 1, 1, 3)
    nrm = torch.where(((view - pos) * fn).sum(-1, keepdim=True) > 0, fn, -fn)
    ro = pos + nrm * 0.001
    kd = torch.rand(B, H, W, 3, generator=g)
    ks = torch.stack([torch.zeros(B, H, W), 0.4 + 0.5 * torch.rand(B, H, W, generator=g), torch.rand(B, H, W, generator=g)], -1)
    mask = torch.ones(B, H, W)
    light = torch.rand(16, 32, 3, generator=g) + 0.2
    pdf, rows, cols = so.light_pdf_tables(light)
    perms = torch.argsort(torch.rand(32768, n * n, generator=g), dim=-1).int()
    od, os_ = so.env_shade(mask, ro, pos, nrm, view, kd, ks, light, pdf, rows, cols, perms, bsdf=0, n_samples_x=n, rnd_seed=7,
                           shadow_scale=1.0, visibility=_brute_force_visibility(va_c, fa_c))
    od0, _ = so.env_shade(mask, ro, pos, nrm, view, kd, ks, light, pdf, rows, cols, perms, bsdf=0, n_samples_x=n, rnd_seed=7,
                          shadow_scale=0.0)
    ctx = ou.OptiXContext()
    ou.optix_build_bvh(ctx, va, fa, rebuild=1)
    gd, gs = ou.optix_env_shade(ctx, mask.to(d), ro.to(d), pos.to(d), nrm.to(d), view.to(d), kd.to(d), ks.to(d), light.to(d),
                                pdf.to(d), rows.to(d), cols.to(d), BSDF="pbr", n_samples_x=n, rnd_seed=7, shadow_scale=1.0,
                                perms=perms.to(d))
    shadowed = float(((od0 - od).abs().sum(-1) > 1e-6).float().mean())
    assert shadowed > 0.3, f"test scene casts too few shadows ({shadowed})"
    # a ray grazing a triangle edge may be classified differently by fp32 CUDA vs torch CPU arithmetic: bound the outliers
    for got, want in ((gd, od), (gs, os_)):
        rel = (got.cpu() - want).abs() / want.abs().clamp(min=1e-3)
        assert float(rel.median()) < 1e-5 and float((rel > 1e-4).float().mean()) < 0.03, (float(rel.median()), float((rel > 1e-4).float().mean()))


def test_shadow_chunking_and_replay_consistent():
    """The wavefront path processes sample pairs in chunks sized by an HBM budget (PCG restarted with an LCG skip-ahead per
    chunk) and the backward pass replays the forward's visibility bits.  Results must not depend on the chunking, and the
    replayed backward must equal a backward that traces again (decorrelated mode with the same seed forced)."""
    import gshell_b200.render.optixutils as ou
    from gshell_b200.render.optixutils import ops
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.grids import bcc_tet_grid
    from oracle import shade_oracle as so
    d = dev()
    v, t = bcc_tet_grid(7)
    g = torch.Generator().manual_seed(9)
    pos3 = (torch.tensor(v) - 0.5) * 2
    sdf = pos3.norm(dim=1) - 0.6 + 0.25 * (torch.rand(v.shape[0], generator=g) - 0.5)
    msdf = torch.rand(v.shape[0], generator=g) - 0.2
    va, fa, _, _, _, _ = GShell_Tets(index_dtype=torch.int32, with_tangents=False)(pos3.to(d), sdf.to(d), msdf.to(d), torch.tensor(t).to(d))
    B, H, W, n = 1, 24, 24, 6                      # 36 sample pairs
    sel = torch.randint(0, fa.shape[0], (B * H * W,), generator=g)
    bary = torch.rand(B * H * W, 3, generator=g); bary = bary / bary.sum(-1, keepdim=True)
    tri = va.cpu()[fa.cpu().long()[sel]]
    pos = (tri * bary[..., None]).sum(1).view(B, H, W, 3)
    fn = torch.nn.functional.normalize(torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), dim=-1).view(B, H, W, 3)
    view = torch.tensor([0.0, 0.0, 3.0]).view(1, 1, 1, 3)
    nrm = torch.where(((view - pos) * fn).sum(-1, keepdim=True) > 0, fn, -fn)
    kd = torch.rand(B, H, W, 3, generator=g)
    ks = torch.stack([torch.zeros(B, H, W), 0.4 + 0.5 * torch.rand(B, H, W, generator=g), torch.rand(B, H, W, generator=g)], -1)
    light = torch.rand(16, 32, 3, generator=g) + 0.2
    pdf, rows, cols = so.light_pdf_tables(light)
    ctx = ou.OptiXContext()
    ou.optix_build_bvh(ctx, va, fa, rebuild=1)

    def run(budget, mask=None):
        mask = torch.ones(B, H, W) if mask is None else mask
        old = ops.SHADOW_SCRATCH_BUDGET
        ops.SHADOW_SCRATCH_BUDGET = budget
        ops._scratch_cache.clear()
        try:
            leaves = [x.clone().to(d).requires_grad_() for x in (pos, nrm, kd, ks, light)]
            dd, ss = ou.optix_env_shade(ctx, mask.to(d), (leaves[0] + 0.001 * leaves[1]).detach(), leaves[0], leaves[1],
                                        view.to(d), leaves[2], leaves[3], leaves[4], pdf.to(d), rows.to(d), cols.to(d), BSDF="pbr",
                                        n_samples_x=n, rnd_seed=3, shadow_scale=1.0)
            (dd.sum() * 0.7 + ss.sum()).backward()
            return dd.detach(), ss.detach(), [x.grad.clone() for x in leaves]
        finally:
            ops.SHADOW_SCRATCH_BUDGET = old
            ops._scratch_cache.clear()
    npix = B * H * W
    d1, s1, g1 = run(1 << 34)                                   # everything in one chunk
    d2, s2, g2 = run(256 + npix * 2 * 33 * 16 + 64)             # 16 pairs per chunk -> 3 chunks (16, 16, 4)
    assert torch.allclose(d1, d2, rtol=1e-5, atol=1e-7) and torch.allclose(s1, s2, rtol=1e-5, atol=1e-7)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(a.abs().max()))
    # sparse coverage: the ray list is sized for the unmasked pixels only (a third of the frame here), so the same budget
    # holds more pairs per chunk; masked pixels stay zero and the covered ones must not change
    sparse = (torch.rand(B, H, W, generator=g) < 0.33).float()
    ncov = int(sparse.sum())
    d3, s3, g3 = run(1 << 34, sparse)
    d4, s4, g4 = run(256 + (ncov * 2 * 32 + npix * 2) * 16 + 64, sparse)
    m = sparse.to(d)[..., None]
    assert torch.allclose(d3, d1 * m, rtol=1e-5, atol=1e-7) and torch.allclose(s3, s1 * m, rtol=1e-5, atol=1e-7)
    assert torch.allclose(d3, d4, rtol=1e-5, atol=1e-7) and torch.allclose(s3, s4, rtol=1e-5, atol=1e-7)
    for a, b in zip(g3, g4):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(a.abs().max()))
    assert float((d1 - so.env_shade(torch.ones(B, H, W), pos + 0.001 * nrm, pos, nrm, view, kd, ks, light, pdf, rows, cols,
                                    ops._EnvShade.perms(n, d).cpu(), bsdf=0, n_samples_x=n, rnd_seed=3, shadow_scale=0.0)[0].to(d)).abs().max()) > 1e-3, \
        "scene casts no shadows: the test would not exercise the visibility path"
