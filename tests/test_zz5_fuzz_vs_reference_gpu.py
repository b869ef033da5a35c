"""Randomised parity of the extraction kernels against the UNMODIFIED reference itself (not its goldens, not the oracle): where the
reference checkout is present (the build container), `GShell_Tets.__call__` (geometry/gshell_tets.py:245-443) and
`GShellFlexiCubes.__call__` (geometry/gshell_flexicubes.py:136-230) are imported through tests/golden/_ref_shim.py and run on the
CPU on small random grids with the inputs that break sign logic -- exact zeros, negative zeros, repeated and tiny values, tets in
shuffled order with shuffled corners -- and the product's kernels must reproduce them with the bars of tests/test_mt_gpu.py /
tests/test_flex_gpu.py (topology and marching-tets floats bit-exact).  The reference checkout does not travel to the GPU box, so on a
B200 these cases skip; they run on the host build of the kernels (GSB_HOST_EMULATION=1, tests/test_emulated_gpu_suite_cpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from _ref_shim import REFERENCE_ROOT, reference_on_cpu      # noqa: E402
sys.path.remove(os.path.join(HERE, "golden"))
from test_flex_gpu import _check as check_flex              # noqa: E402
from test_mt_gpu import _check_forward, _check_grads, _run_cuda      # noqa: E402

needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="needs the reference checkout (build container)")


def _special(values, gen, kind):
    """inject the values that decide sign tests"""
    n = values.shape[0]
    r = torch.rand(n, generator=gen)
    if kind == "zeros":
        values[r < 0.25] = 0.0
        values[(r >= 0.25) & (r < 0.35)] = -0.0
    elif kind == "quantised":
        values = torch.round(values * 4) / 4                      # many equal values, many exact zeros
    elif kind == "tiny":
        values[r < 0.3] *= 1e-30
    return values


@needs_reference
@pytest.mark.parametrize("seed", range(64))
def test_marching_tets_matches_the_unmodified_reference(seed):
    from gshell_b200.grids import bcc_tet_grid
    gen = torch.Generator().manual_seed(1000 + seed)
    n = [2, 3, 4, 5][seed % 4]
    kind = ["plain", "zeros", "quantised", "tiny"][(seed // 4) % 4]
    v, t = bcc_tet_grid(n)
    pos = (torch.tensor(v) - 0.5 + 0.05 * (torch.rand(v.shape, generator=gen) - 0.5)).float()
    nv = v.shape[0]
    sdf = _special(torch.rand(nv, generator=gen) - 0.35, gen, kind).float()
    msdf = torch.rand(nv, generator=gen) - 0.4
    msdf = _special(torch.where(torch.rand(nv, generator=gen) < 0.5, msdf, -msdf), gen, kind).float()
    tets = torch.tensor(t)
    if seed % 16 >= 8:                                              # any tet list, not only the generator's order
        tets = tets[torch.randperm(tets.shape[0], generator=gen)]
        tets = torch.stack([row[torch.randperm(4, generator=gen)] for row in tets])
    leaves_ref = [x.clone().requires_grad_() for x in (pos, sdf, msdf)]
    with reference_on_cpu() as imp:
        ref = imp("geometry.gshell_tets").GShell_Tets()
        va, fa, _, _, _, extra = ref(*leaves_ref, tets)
        gw = torch.Generator().manual_seed(seed)
        w = {"wa": torch.randn(va.shape, generator=gw), "wm": torch.randn(extra["msdf"].shape, generator=gw),
             "ww": torch.randn(extra["vertices_watertight"].shape, generator=gw)}
        want_grads = None
        if va.shape[0]:
            probe = (va * w["wa"]).sum() + (extra["msdf"] * w["wm"]).sum() + (extra["vertices_watertight"] * w["ww"]).sum()
            want_grads = [torch.zeros_like(x) if g is None else g for g, x in
                          zip(torch.autograd.grad(probe, leaves_ref, allow_unused=True), leaves_ref)]
    want = {"faces_aug": fa, "faces_watertight": extra["faces_watertight"], "n_verts_watertight": extra["n_verts_watertight"],
            "verts_aug": va.detach(), "vertices_watertight": extra["vertices_watertight"].detach(), "msdf_aug": extra["msdf"].detach(),
            "msdf_watertight": extra["msdf_watertight"].detach(), "msdf_boundary": extra["msdf_boundary"].detach()}
    leaves, out = _run_cuda(pos, sdf, msdf, tets)
    _check_forward(out, want)
    if want_grads is not None:
        _check_grads(leaves, out, w, want_grads)


@needs_reference
@pytest.mark.parametrize("seed", range(24))
def test_flexicubes_matches_the_unmodified_reference(seed):
    from gshell_b200.geometry.gshell_flexicubes import GShellFlexiCubes
    gen = torch.Generator().manual_seed(2000 + seed)
    res = [3, 4, 5, 6][seed % 4]
    kind = ["plain", "zeros", "quantised"][(seed // 4) % 3]
    d = device()
    fc = GShellFlexiCubes(device=d)
    verts, cubes = fc.construct_voxel_grid(res)
    verts, cubes = verts.cpu(), cubes.cpu()
    nv, nc = verts.shape[0], cubes.shape[0]
    x = (verts + 0.2 / res * (torch.rand(nv, 3, generator=gen) - 0.5)).float()
    s = _special(verts.norm(dim=1) - 0.35 + 0.2 * (torch.rand(nv, generator=gen) - 0.5), gen, kind).float()
    nu = _special(verts[:, 1] + 0.1 + 0.3 * (torch.rand(nv, generator=gen) - 0.5), gen, kind).float()
    wgt = (torch.randn(nc, 21, generator=gen) * 0.5).float()
    lr = [t.clone().requires_grad_() for t in (x, s, nu, wgt)]
    with reference_on_cpu() as imp:
        ref = imp("geometry.gshell_flexicubes").GShellFlexiCubes(device="cpu")
        rverts, rcubes = ref.construct_voxel_grid(res)
        assert torch.equal(rcubes, cubes) and torch.equal(rverts, verts)
        vo, fa, L, ex = ref(lr[0], lr[1], lr[2], cubes, res, lr[3][:, :12], lr[3][:, 12:20], lr[3][:, 20])
        gw = torch.Generator().manual_seed(seed)
        w = {"wv": torch.randn(vo.shape, generator=gw), "wm": torch.randn(ex["msdf"].shape, generator=gw),
             "wl": torch.randn(L.shape, generator=gw), "ww": torch.randn(ex["vertices_watertight"].shape, generator=gw)}
        grads = None
        if vo.shape[0]:
            probe = (vo * w["wv"]).sum() + (ex["msdf"] * w["wm"]).sum() + (L * w["wl"]).sum() + (ex["vertices_watertight"] * w["ww"]).sum()
            grads = [torch.zeros_like(t) if g is None else g for g, t in zip(torch.autograd.grad(probe, lr, allow_unused=True), lr)]
    gl = [t.clone().to(d).requires_grad_() for t in (x, s, nu, wgt)]
    out = fc(gl[0], gl[1], gl[2], cubes.to(d), res, gl[3][:, :12], gl[3][:, 12:20], gl[3][:, 20])
    if vo.shape[0] == 0:
        assert out[0].shape[0] == 0 and out[1].shape == (0, 3)
        return
    want = {"faces_open": fa, "faces_watertight": ex["faces_watertight"], "n_verts_watertight": ex["n_verts_watertight"],
            "vertices_open": vo.detach(), "vertices_watertight": ex["vertices_watertight"].detach(), "msdf": ex["msdf"].detach(),
            "msdf_watertight": ex["msdf_watertight"].detach(), "msdf_boundary": ex["msdf_boundary"].detach(), "L_dev": L.detach()}
    check_flex(out, want, gl, w, grads)


def _reference_function(rel_path, name):
    """one top-level function of a reference module that cannot be imported here (it pulls OptiX / kaolin): compiled from its own
    source lines at run time, nothing copied"""
    import ast
    src = open(os.path.join(REFERENCE_ROOT, rel_path)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), rel_path, "exec"), ns)
    return ns[name]


@needs_reference
@pytest.mark.parametrize("seed", range(8))
def test_sdf_regulariser_matches_the_unmodified_reference(seed):
    """compute_sdf_reg_loss (geometry/gshell_tets_geometry.py:33-39) on random edge lists, SDF with exact zeros (sign(0) = 0
    differs from both signs) and with no sign change at all."""
    from gshell_b200.geometry.gshell_tets_geometry import compute_sdf_reg_loss
    gen = torch.Generator().manual_seed(3000 + seed)
    nv = 50 + 137 * seed
    sdf = (torch.rand(nv, generator=gen) - (0.35 if seed != 7 else -0.1)).float()
    sdf = _special(sdf, gen, ["plain", "zeros", "quantised", "tiny"][seed % 4]).float()
    edges = torch.randint(0, nv, (5 * nv, 2), generator=gen)
    ref_fn = _reference_function("geometry/gshell_tets_geometry.py", "compute_sdf_reg_loss")
    a = sdf.clone().requires_grad_()
    want = ref_fn(a, edges)
    b = sdf.clone().to(device()).requires_grad_()
    got = compute_sdf_reg_loss(b, edges.to(device()))
    if not bool(torch.isfinite(want)):          # no edge with a sign change: the reference's mean over nothing
        assert not bool(torch.isfinite(got.cpu())) or float(got) == 0.0
        return
    assert abs(float(got) - float(want)) <= 1e-5 * max(abs(float(want)), 1e-6), (float(got), float(want))
    gw, = torch.autograd.grad(want, a)
    gg, = torch.autograd.grad(got, b)
    assert float((gg.cpu() - gw).abs().max()) <= 1e-5 * float(gw.abs().max().clamp(min=1e-12))


@needs_reference
@pytest.mark.parametrize("seed", range(6))
def test_light_tables_match_the_unmodified_reference(seed):
    """EnvironmentLight.update_pdf (render/light.py:46-59) on random probes, incl. one with black rows (zero column sums)."""
    import types
    from gshell_b200.render import light
    gen = torch.Generator().manual_seed(4000 + seed)
    h, w = [(16, 16), (16, 32), (32, 64), (64, 64), (256, 256), (48, 16)][seed]
    base = (torch.rand(h, w, 3, generator=gen) * 2.0 + 0.01).float()
    if seed % 2:
        base[h // 3] = 0.0
        base[:, w // 2] = 0.0
    sys.modules.setdefault("tinycudann", types.ModuleType("tinycudann"))
    with reference_on_cpu() as imp:
        ref = imp("render.light").EnvironmentLight(base.clone())
        want = {"pdf": ref._pdf.clone(), "cols": ref.cols.clone(), "rows": ref.rows.clone()}
    lgt = light.EnvironmentLight(base.clone().to(device()))
    for name, got in (("pdf", lgt._pdf), ("cols", lgt.cols), ("rows", lgt.rows)):
        ref_t = want[name]
        assert got.shape == ref_t.shape, name
        assert float((got.cpu() - ref_t).abs().max()) <= 2e-6 * float(ref_t.abs().max()) + 1e-7, (name, float((got.cpu() - ref_t).abs().max()))


def _rel_close(got, want, tol, what):
    got, want = got.detach().cpu().float(), want.detach().float()
    scale = float(want.abs().max().clamp(min=1e-6))
    err = float((got - want).abs().max())
    assert got.shape == want.shape and err <= tol * scale, (what, err, scale)


@needs_reference
@pytest.mark.parametrize("seed", range(8))
def test_gbuffer_operators_match_the_unmodified_reference(seed):
    """prepare_shading_normal / image_loss / xfm_points against the reference's own PyTorch statements of them (`use_python=True`,
    render/renderutils/ops.py:197-236, 479-503, 518-533) on random shapes, flags and broadcast operands."""
    import types
    import gshell_b200.render.renderutils as ru
    gen = torch.Generator().manual_seed(5000 + seed)
    N = lambda *s: torch.randn(*s, generator=gen)          # noqa: E731
    R = lambda *s: torch.rand(*s, generator=gen)           # noqa: E731
    # no extent of 3 before the channel axis: the reference's Python twin calls torch.cross without `dim` (bsdf.py:40), which then
    # picks the FIRST axis of size 3 -- its CUDA kernel, the contract, always crosses per pixel
    B, H, W = [1, 2, 4][seed % 3], 5 + 3 * seed, 4 + 5 * (seed % 4)
    two_sided, opengl, with_pert = bool(seed & 1), bool(seed & 2), bool(seed & 4)
    ins = {"pos": N(B, H, W, 3), "view_pos": N(B, 1, 1, 3) * 3, "smooth_nrm": N(B, H, W, 3), "smooth_tng": N(B, H, W, 3),
           "geom_nrm": torch.nn.functional.normalize(N(B, H, W, 3), dim=-1)}
    pert = torch.nn.functional.normalize(N(B, H, W, 3), dim=-1) * torch.tensor([0.3, 0.3, 1.0]) if with_pert else None
    img, tgt = R(B, H, W, 3) * 3.0, R(B, H, W, 3) * 3.0
    pts, mtx = N(1, 50 + seed, 3), N(B, 4, 4)
    sys.modules.setdefault("tinycudann", types.ModuleType("tinycudann"))
    want = {}
    with reference_on_cpu() as imp:
        rru = imp("render.renderutils")
        lm = imp("render.renderutils.loss")
        la = {k: v.clone().requires_grad_() for k, v in ins.items()}
        lp = pert.clone().requires_grad_() if with_pert else None
        out = rru.prepare_shading_normal(la["pos"], la["view_pos"], lp, la["smooth_nrm"], la["smooth_tng"], la["geom_nrm"],
                                         two_sided_shading=two_sided, opengl=opengl, use_python=True)
        wn = N(*out.shape)
        leaves = list(la.values()) + ([lp] if with_pert else [])
        want["nrm"] = (out.detach(), wn, torch.autograd.grad((out * wn).sum(), leaves))
        for loss in ("l1", "mse", "smape", "relmse"):
            for tm in ("none", "log_srgb"):
                a, b = img.clone().requires_grad_(), tgt.clone().requires_grad_()
                if tm == "log_srgb":        # the reference's CUDA kernel applies no exposure factor (loss.cu:38-41); its Python twin does (loss.py:16-18)
                    val = lm.image_loss_fn(lm._tonemap_srgb(torch.log(torch.clamp(a, min=0, max=65535) + 1), exposure=1),
                                           lm._tonemap_srgb(torch.log(torch.clamp(b, min=0, max=65535) + 1), exposure=1), loss, "none")
                else:
                    val = rru.image_loss(a, b, loss=loss, tonemapper=tm, use_python=True)
                want[loss, tm] = (val.detach(), torch.autograd.grad(val, [a, b]))
        p = pts.clone().requires_grad_()
        o = rru.xfm_points(p, mtx, use_python=True)
        wx = N(*o.shape)
        want["xfm"] = (o.detach(), wx, torch.autograd.grad((o * wx).sum(), p)[0])
    d = device()
    ga = {k: v.clone().to(d).requires_grad_() for k, v in ins.items()}
    gp = pert.clone().to(d).requires_grad_() if with_pert else None
    out = ru.prepare_shading_normal(ga["pos"], ga["view_pos"], gp, ga["smooth_nrm"], ga["smooth_tng"], ga["geom_nrm"],
                                    two_sided_shading=two_sided, opengl=opengl)
    _rel_close(out, want["nrm"][0], 1e-4, "shading normal")
    grads = torch.autograd.grad((out * want["nrm"][1].to(d)).sum(), list(ga.values()) + ([gp] if with_pert else []))
    for name, g, w in zip(list(ins) + ["perturbed_nrm"], grads, want["nrm"][2]):
        _rel_close(g, w, 1e-4, "shading normal d/d " + name)
    for loss in ("l1", "mse", "smape", "relmse"):
        for tm in ("none", "log_srgb"):
            a, b = img.clone().to(d).requires_grad_(), tgt.clone().to(d).requires_grad_()
            val = ru.image_loss(a, b, loss=loss, tonemapper=tm)
            _rel_close(val, want[loss, tm][0], 1e-5, f"image_loss {loss}/{tm}")
            for g, w in zip(torch.autograd.grad(val, [a, b]), want[loss, tm][1]):
                _rel_close(g, w, 1e-4, f"image_loss {loss}/{tm} gradient")
    p = pts.clone().to(d).requires_grad_()
    o = ru.xfm_points(p, mtx.to(d))
    _rel_close(o, want["xfm"][0], 1e-5, "xfm_points")
    _rel_close(torch.autograd.grad((o * want["xfm"][1].to(d)).sum(), p)[0], want["xfm"][2], 1e-5, "xfm_points gradient")


@needs_reference
@pytest.mark.parametrize("seed", range(6))
def test_denoiser_matches_the_reference_python_filter(seed):
    """bilateral_denoiser against the Python BilateralDenoiser the reference keeps beside its kernel
    (render/optixutils/tests/filter_test.py:31-74), several filter widths and image shapes."""
    import math
    import gshell_b200.render.optixutils as ou
    gen = torch.Generator().manual_seed(6000 + seed)
    N = lambda *s: torch.randn(*s, generator=gen)          # noqa: E731
    R = lambda *s: torch.rand(*s, generator=gen)           # noqa: E731
    sigma = [0.3, 0.6, 1.0, 1.7, 2.4, 0.05][seed]
    B, H, W = 1 + seed % 2, 9 + 4 * seed, 23 - 3 * seed
    src = open(os.path.join(REFERENCE_ROOT, "render/optixutils/tests/filter_test.py")).read()
    ns = {"torch": torch, "np": np, "math": math, "dot": lambda a, b: torch.sum(a * b, -1, keepdim=True)}
    exec(src[src.index("class BilateralDenoiser"):src.index("def relative_loss")], ns)
    inp = R(B, H, W, 11)
    inp[..., 3:6] = torch.nn.functional.normalize(N(B, H, W, 3) * 0.3 + torch.tensor([0.0, 0.0, 1.0]), dim=-1)
    inp[..., 9] = 0.9 + 0.05 * R(B, H, W)
    inp[..., 10] = 0.001 + 0.004 * R(B, H, W)
    a = inp.clone().requires_grad_()
    with reference_on_cpu():                                   # the filter allocates on 'cuda'
        want = ns["BilateralDenoiser"](sigma=sigma).forward(a)
        w = N(*want.shape)
        g_want = torch.autograd.grad((want * w).sum(), a)[0][..., 0:3]
    d = device()
    col = inp[..., 0:3].clone().to(d).requires_grad_()
    out = ou.bilateral_denoiser(col, inp[..., 3:6].contiguous().to(d), inp[..., 9:11].contiguous().to(d), sigma)
    _rel_close(out, want, 1e-4, "denoiser")
    _rel_close(torch.autograd.grad((out * w.to(d)).sum(), col)[0], g_want, 1e-4, "denoiser gradient")


@needs_reference
@pytest.mark.parametrize("seed", range(12))
def test_generative_decode_matches_the_unmodified_reference(seed):
    """GShell_Tets.marching_from_auggrid (geometry/gshell_tets.py:446-629) on random augmented grids (construction as in
    tests/golden/make_golden_auggrid.py): topology, tet ids bit-exact; positions and mSDF to fp32 rounding.  (The tangent output has
    rows that no input determines -- cancelling face normals; it is compared row-wise in tests/test_zz3_generative_decode_gpu.py.)"""
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.grids import bcc_tet_grid
    gen = torch.Generator().manual_seed(7000 + seed)
    n = [2, 3, 4][seed % 3]
    v, t = bcc_tet_grid(n)
    verts = torch.tensor(v, dtype=torch.float32)
    tets = torch.tensor(t, dtype=torch.long)
    disc = torch.round(verts * (4 * n)).long().float()
    pos = (verts - 0.5) * 2.0 + 0.2 / n * (torch.rand(verts.shape, generator=gen) - 0.5)
    frac = [0.2, 0.35, 0.5, 0.65][seed % 4]
    sdf = torch.sign(torch.rand(verts.shape[0], generator=gen) - frac) if seed % 5 else torch.sign((verts - 0.5).norm(dim=1) - 0.3)
    sdf[sdf == 0] = 1.0
    sorted_edges = torch.sort(tets[:, [0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3]].reshape(-1, 6, 2), dim=-1)[0]
    G = 4 * n + 1
    coeff = torch.rand(G, G, G, generator=gen) * 1.6 - 0.3
    msdf_sign = torch.sign(torch.rand(G, G, G, generator=gen) - [0.2, 0.5, 0.8][seed % 3])
    occ = torch.rand(8 * n + 1, 8 * n + 1, 8 * n + 1, generator=gen) * 2 - 1
    with reference_on_cpu() as imp:
        ref = imp("geometry.gshell_tets").GShell_Tets()
        rva, rfa, _, _, _, rv, rgidx, rm_aug, rm = ref.marching_from_auggrid(pos, sdf, tets, sorted_edges, coeff, disc, msdf_sign, occ)
    d = device()
    va, fa, a, b, tng, vv, gidx, m_aug, m = GShell_Tets().marching_from_auggrid(
        pos.to(d), sdf.to(d), tets.to(d), sorted_edges.to(d), coeff.to(d), disc.to(d), msdf_sign.to(d), occ.to(d))
    assert torch.equal(fa.cpu().long(), rfa.long()) and torch.equal(gidx.cpu().long(), rgidx.long())
    for name, got, want in (("verts_aug", va, rva), ("verts", vv, rv), ("msdf_aug", m_aug, rm_aug), ("msdf", m, rm)):
        assert got.shape == want.shape, name
        if got.numel():
            assert float((got.cpu() - want).abs().max()) <= 1e-5, name
    assert tng.shape == rva.shape


@needs_reference
@pytest.mark.parametrize("seed", range(6))
def test_regularisers_match_the_unmodified_reference(seed):
    """chroma_loss / shading_loss / material_smoothness_grad (render/regularizer.py:21-52) on random buffers of odd sizes: values and
    gradients of a weighted sum against the reference module's."""
    import types
    from gshell_b200.render import regularizer as reg
    gen = torch.Generator().manual_seed(8000 + seed)
    R = lambda *s: torch.rand(*s, generator=gen)           # noqa: E731
    B, H, W = 1 + seed % 3, 7 + 5 * seed, 11 + 3 * (seed % 4)
    alpha = (R(B, H, W, 1) > 0.3).float()
    base = {"diff": R(B, H, W, 3) * 2, "spec": R(B, H, W, 3), "kd": R(B, H, W, 3), "kd_grad": torch.cat([R(B, H, W, 3), alpha], -1),
            "ks_grad": torch.cat([R(B, H, W, 3) * torch.tensor([0.0, 1.0, 1.0]), alpha], -1), "nrm_grad": torch.cat([R(B, H, W, 3), alpha], -1)}
    color_ref = torch.cat([R(B, H, W, 3), alpha], -1)
    lam = dict(diffuse=0.15, specular=0.0025, chroma=0.3, kd=0.25, ks=0.1, nrm=0.05 + 0.1 * seed)

    def total(mod, t, ref_dev):
        l_sh = mod.shading_loss(t["diff"], t["spec"], ref_dev, lam["diffuse"], lam["specular"])
        l_ch = mod.chroma_loss(t["kd"], ref_dev, lam["chroma"])
        l_ms = mod.material_smoothness_grad(t["kd_grad"], t["ks_grad"], t["nrm_grad"], lambda_kd=lam["kd"], lambda_ks=lam["ks"], lambda_nrm=lam["nrm"])
        return l_sh, l_ch, l_ms
    sys.modules.setdefault("tinycudann", types.ModuleType("tinycudann"))
    with reference_on_cpu() as imp:
        rmod = imp("render.regularizer")
        rt = {k: v.clone().requires_grad_() for k, v in base.items()}
        rl = total(rmod, rt, color_ref)
        rg = torch.autograd.grad(rl[0] + 2.0 * rl[1] + 3.0 * rl[2], list(rt.values()))
    d = device()
    gt = {k: v.clone().to(d).requires_grad_() for k, v in base.items()}
    gl = total(reg, gt, color_ref.to(d))
    for name, a, b in zip(("shading", "chroma", "smoothness"), gl, rl):
        assert abs(float(a) - float(b)) <= 1e-5 * max(abs(float(b)), 1e-6), (name, float(a), float(b))
    gg = torch.autograd.grad(gl[0] + 2.0 * gl[1] + 3.0 * gl[2], list(gt.values()))
    for name, a, b in zip(base, gg, rg):
        if name.endswith("_grad"):       # channel 3 of these buffers is the coverage mask: no gradient path in the renderer, the kernel writes 0
            a, b = a[..., :3], b[..., :3]
        _rel_close(a, b, 1e-4, "d/d " + name)


@needs_reference
@pytest.mark.parametrize("seed", range(6))
def test_vertex_normals_match_the_unmodified_reference(seed):
    """mesh.auto_normals (render/mesh.py:212-237) on random triangle soups with shared vertices, degenerate faces (repeated
    corners, collinear corners) and vertices no face references (the reference's (0, 0, 1) default): values and the gradient to
    the positions."""
    import types
    from gshell_b200.render import mesh
    gen = torch.Generator().manual_seed(9000 + seed)
    nv, nf = 40 + 30 * seed, 90 + 70 * seed
    v = torch.randn(nv, 3, generator=gen)
    f = torch.randint(0, nv - 5, (nf, 3), generator=gen)           # the last 5 vertices stay unreferenced
    f[::11, 1] = f[::11, 0]                                        # degenerate: a repeated corner
    v[f[5]] = v[f[5, 0]] + torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0]])     # degenerate: collinear
    w = torch.randn(nv, 3, generator=gen)
    sys.modules.setdefault("tinycudann", types.ModuleType("tinycudann"))
    with reference_on_cpu() as imp:
        rmesh = imp("render.mesh")
        a = v.clone().requires_grad_()
        want = rmesh.auto_normals(rmesh.Mesh(a, f)).v_nrm
        g_want, = torch.autograd.grad((want * w).sum(), a)
    d = device()
    b = v.clone().to(d).requires_grad_()
    got = mesh.auto_normals(mesh.Mesh(b, f.to(d))).v_nrm
    # a vertex whose only faces are degenerate accumulates rounding residue (|sum|^2 ~ 1e-17 here): the reference normalises that
    # residue into an arbitrary direction, a differently rounded cross product lands on the (0, 0, 1) default -- neither is
    # determined by the input; compare the rows that are
    fn = torch.linalg.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    acc = torch.zeros_like(v)
    for k in range(3):
        acc.index_add_(0, f[:, k], fn)
    det = (acc * acc).sum(-1) > 1e-10
    det[-5:] = True
    assert int((~det).sum()) <= 3 and bool(torch.isfinite(got).all())
    _rel_close(got.cpu()[det], want[det], 2e-5, "vertex normals")
    g_got, = torch.autograd.grad((got * (w * det[:, None]).to(d)).sum(), b)
    with reference_on_cpu() as imp:
        rmesh = imp("render.mesh")
        a2 = v.clone().requires_grad_()
        g_want, = torch.autograd.grad((rmesh.auto_normals(rmesh.Mesh(a2, f)).v_nrm * (w * det[:, None])).sum(), a2)
    _rel_close(g_got, g_want, 1e-4, "vertex normals gradient")
    assert bool((got[-5:].cpu() == torch.tensor([0.0, 0.0, 1.0])).all())


@pytest.mark.skipif(DEVICE != "cpu", reason="bounds calibrated for the host build (IEEE arithmetic, no fast math); the device run of this comparison "
                    "with its measured bounds is tests/test_shade_gpu.py::test_env_shade_vs_compiled_reference_at_benchmark_sample_counts")
@pytest.mark.parametrize("seed", range(9))
def test_integrator_matches_the_reference_kernel_compiled_for_the_cpu(seed):
    """optix_env_shade against the reference's own envsampling/kernel.cu (oracle/_ref, compiled unmodified for the CPU; prebuilt, so
    no checkout is needed): the three BSDF modes, n = 1..3, with and without occluders, random G-buffers at roughness >= 0.3 (below
    that the reference's own result depends on its compiler flags, tests/test_oracle_env_shade_conditioning.py), forward and the
    five gradients."""
    import gshell_b200.render.optixutils as ou
    from oracle import ref_env_shade as ref, shade_oracle as so
    from test_shade_gpu import _shade_inputs
    try:
        ref.lib()
    except Exception as e:                         # noqa: BLE001
        pytest.skip(f"oracle/_ref not built: {e}")
    bsdf = seed % 3
    n = 1 + (seed // 3) % 3
    shadows = seed % 2 == 0
    B, H, W = 1 + seed % 2, 12 + seed, 20 - seed
    mask, pos, nrm, view, kd, ks, light = _shade_inputs(B, H, W, 700 + seed, rough_min=0.3, lh=16 * (1 + seed % 2), lw=32)
    g = torch.Generator().manual_seed(seed)
    pdf, rows, cols = so.light_pdf_tables(light)
    perms = torch.argsort(torch.rand(32768, n * n, generator=g), dim=-1).int()
    d = device()
    ctx = ou.OptiXContext()
    verts = tris = None
    if shadows:
        c = torch.randn(120, 1, 3, generator=g) * 0.5
        verts = (c + 0.15 * torch.randn(120, 3, 3, generator=g)).reshape(-1, 3)
        tris = torch.arange(360, dtype=torch.int32).reshape(-1, 3)
        ou.optix_build_bvh(ctx, verts.to(d), tris.to(d), rebuild=1)
    ss = 1.0 if shadows else 0.0
    ro = pos + 0.001 * nrm
    a = (mask, ro, pos, nrm, view, kd, ks, light, pdf, rows, cols, perms)
    rd, rs = ref.env_shade_fwd(*a, bsdf=bsdf, n_samples_x=n, rnd_seed=11 + seed, shadow_scale=ss, verts=verts, tris=tris)
    gen = torch.Generator().manual_seed(99)
    wd, ws = torch.randn(rd.shape, generator=gen), torch.randn(rs.shape, generator=gen)
    rgrads = ref.env_shade_bwd(*a, wd, ws, bsdf=bsdf, n_samples_x=n, rnd_seed=11 + seed, shadow_scale=ss, verts=verts, tris=tris)
    gl = [t.clone().to(d).requires_grad_() for t in (pos, nrm, kd, ks, light)]
    gd, gs = ou.optix_env_shade(ctx, mask.to(d), ro.to(d), gl[0], gl[1], view.to(d), gl[2], gl[3], gl[4], pdf.to(d), rows.to(d), cols.to(d),
                                BSDF=["pbr", "diffuse", "white"][bsdf], n_samples_x=n, rnd_seed=11 + seed, shadow_scale=ss, perms=perms.to(d))
    cov = mask > 0
    for name, got, want in (("diff", gd, rd), ("spec", gs, rs)):
        got = got.detach().cpu()
        floor = 1e-3 * want[cov].abs().mean().clamp(min=1e-8)
        rel = ((got - want).abs() / want.abs().clamp(min=floor))[cov]
        # a ray grazing an occluder's edge, or a sample on a texel border of the probe, may fall on the other side in the other
        # implementation: a bounded share of the pixels, everything else to fp32 rounding
        assert float(rel.median()) < 1e-5 and float((rel > 1e-4).float().mean()) < 0.02, (name, float(rel.median()), float((rel > 1e-4).float().mean()))
    ((gd * wd.to(d)).sum() + (gs * ws.to(d)).sum()).backward()
    for name, x, want in zip(("pos", "nrm", "kd", "ks", "light"), gl, rgrads):
        if float(want.norm()) == 0.0:
            assert float(x.grad.norm()) == 0.0, name
            continue
        l2 = float((x.grad.cpu() - want).norm() / want.norm())
        assert l2 < 5e-3, (name, l2)
