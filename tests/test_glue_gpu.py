"""Loss assembly of tick() on the CUDA reduction kernels (csrc/tick_ops.cu), against the
unmodified reference's values and gradients (tests/golden/glue_losses.npz, shade_light.npz; generators beside them) and, for
the terms the reference writes inline in tick() (gshell_tets_geometry.py:283-290, 325-356), against the same torch expressions
restated here."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _device import DEVICE               # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
D = DEVICE


def _load(name):
    z = np.load(os.path.join(HERE, "golden", name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _close(a, b, tol=1e-5):
    a, b = a.detach().cpu().float(), b.float()
    return torch.allclose(a, b, rtol=tol, atol=tol * float(b.abs().max()) + 1e-12)


def test_regularizers_match_reference():
    from gshell_b200.render import regularizer as reg
    g = _load("glue_losses.npz")
    leaves = {k: g[k].clone().to(D).requires_grad_() for k in ("diff", "spec", "kd", "kd_grad", "ks_grad", "nrm_grad")}
    ref = g["color_ref"].to(D)
    l_sh = reg.shading_loss(leaves["diff"], leaves["spec"], ref, 0.15, 0.0025)
    l_ch = reg.chroma_loss(leaves["kd"], ref, 0.3)
    l_ms = reg.material_smoothness_grad(leaves["kd_grad"], leaves["ks_grad"], leaves["nrm_grad"], lambda_kd=0.25, lambda_ks=0.1,
                                        lambda_nrm=0.05)
    assert _close(l_sh, g["shading_loss"]) and _close(l_ch, g["chroma_loss"]) and _close(l_ms, g["material_smoothness"])
    grads = torch.autograd.grad(l_sh + 2.0 * l_ch + 3.0 * l_ms, list(leaves.values()))
    for k, got in zip(leaves, grads):
        want = g[f"g_{k}"]
        if k in ("kd_grad", "ks_grad", "nrm_grad"):
            # the alpha channel of these buffers is the coverage mask (no gradient path in the renderer); the kernel writes 0 there
            got, want = got[..., :3], want[..., :3]
        assert _close(got, want), k


def test_all_terms_in_one_pass_equal_the_separate_calls():
    from gshell_b200 import losses
    from gshell_b200.render import regularizer as reg
    g = _load("glue_losses.npz")
    t = {k: g[k].to(D) for k in g}
    lam = (0.3, 0.15, 0.0025, 0.25, 0.1, 0.05)
    pad = lambda x: F.pad(x, (0, 1), value=1.0)       # noqa: E731
    shaded = torch.rand(t["diff"].shape[:-1] + (4,), device=D).requires_grad_()
    msdf = (torch.rand(t["diff"].shape[:-1] + (2,), device=D) - 0.5).requires_grad_()
    out = losses.image_terms(t["color_ref"], 31, lam, shaded=shaded, msdf_img=msdf, kd=pad(t["kd"]), kd_grad=t["kd_grad"],
                             ks_grad=t["ks_grad"], nrm_grad=t["nrm_grad"], diffuse=pad(t["diff"]), specular=pad(t["spec"]))
    want_reg = reg.shading_loss(t["diff"], t["spec"], t["color_ref"], lam[1], lam[2]) + reg.chroma_loss(t["kd"], t["color_ref"], lam[0]) + \
        reg.material_smoothness_grad(t["kd_grad"], t["ks_grad"], t["nrm_grad"], lam[3], lam[4], lam[5])
    assert _close(out[1], want_reg.cpu())
    # the inline terms of tick() (reference gshell_tets_geometry.py:283-290), restated with torch
    ref = t["color_ref"]
    gt = ref[..., 3:]
    s2, m2 = shaded.detach().clone().requires_grad_(), msdf.detach().clone().requires_grad_()
    want_img = F.mse_loss(s2[..., 3:], gt) + 0.5 * F.l1_loss(m2.clamp(min=0) * (gt == 0).float(), torch.zeros_like(m2)) + \
        0.5 * F.l1_loss(m2.clamp(max=0) * (gt == 1).float(), torch.ones_like(m2))
    assert _close(out[0], want_img.detach().cpu())
    ga = torch.autograd.grad(out[0], [shaded, msdf])
    gb = torch.autograd.grad(want_img, [s2, m2])
    assert _close(ga[0], gb[0].cpu()) and _close(ga[1], gb[1].cpu())


def test_sdf_regulariser_matches_reference():
    from gshell_b200.geometry.gshell_tets_geometry import compute_sdf_reg_loss
    g = _load("glue_losses.npz")
    sdf = g["sdf"].clone().to(D).requires_grad_()
    loss = compute_sdf_reg_loss(sdf, g["edges"].to(D))
    assert _close(loss, g["sdf_reg"])
    assert _close(torch.autograd.grad(loss, sdf)[0], g["g_sdf"])
    assert int((g["sdf"] == 0).sum()) > 0            # the sign(0) edge case is part of the fixture


def test_sdf_regulariser_full_size_edge_table():
    """BASELINE '256' grid: 15 M static edges; value and gradient against the torch expression of the reference."""
    from gshell_b200 import losses
    from gshell_b200.geometry.tet_tables import tables_for
    from gshell_b200.grids import bcc_tet_grid
    v, t = bcc_tet_grid(103)
    tets = torch.tensor(t).to(D)
    edges = tables_for(tets, v.shape[0]).edge_v
    sdf = (torch.rand(v.shape[0], device=D) - 0.1).requires_grad_()
    loss = losses.sdf_reg_loss(sdf, edges)
    gk = torch.autograd.grad(loss, sdf)[0]
    s2 = sdf.detach().clone().requires_grad_()
    s = s2[edges.reshape(-1).long()].reshape(-1, 2)
    s = s[torch.sign(s[..., 0]) != torch.sign(s[..., 1])]
    want = F.binary_cross_entropy_with_logits(s[..., 0], (s[..., 1] > 0).float()) + \
        F.binary_cross_entropy_with_logits(s[..., 1], (s[..., 0] > 0).float())
    gw = torch.autograd.grad(want, s2)[0]
    assert abs(float(loss) - float(want)) < 1e-5 * abs(float(want))
    assert float((gk - gw).abs().max()) < 1e-5 * float(gw.abs().max())


def test_msdf_regularisers_match_torch_expression():
    from gshell_b200 import losses
    g = torch.Generator().manual_seed(3)
    n_wt, n_b, F_ = 500, 300, 900
    msdf = ((torch.rand(n_wt + n_b, generator=g) - 0.5) * 4).to(D).requires_grad_()       # |m| up to 2: both Huber branches
    tris = torch.randint(0, n_wt + n_b, (F_, 3), generator=g).int().to(D)
    visible = torch.randperm(F_, generator=g)[:200].sort().values.to(D)
    eps, wo, wc = 1e-3, 0.7, 1.3
    mb = msdf[n_wt:]
    bmask = losses.visible_boundary_mask(tris, visible, n_wt, n_b)
    loss = losses.msdf_reg_loss(msdf, mb, bmask, wo, wc, eps)
    got = torch.autograd.grad(loss, msdf)[0]
    # reference expression (gshell_tets_geometry.py:325-356)
    m2 = msdf.detach().clone().requires_grad_()
    e = torch.tensor([eps], device=D)
    vis = tris[visible].reshape(-1).long()
    sel = torch.unique(vis[vis >= n_wt]) - n_wt
    assert torch.equal(bmask.nonzero()[:, 0], sel)
    want = wo * F.huber_loss(m2.clamp(min=-e).squeeze(), -e.expand(m2.size(0)), reduction="sum") + \
        wc * F.huber_loss(m2[n_wt:][sel].clamp(max=e).squeeze(), e.expand(sel.numel()), reduction="sum")
    gw = torch.autograd.grad(want, m2)[0]
    assert abs(float(loss) - float(want)) < 1e-5 * abs(float(want))
    assert float((got - gw).abs().max()) < 1e-5 * float(gw.abs().max())


def test_light_tables_match_reference():
    from gshell_b200.render import light
    g = _load("shade_light.npz")
    lgt = light.EnvironmentLight(g["base"].to(D))
    assert lgt.rows.shape == lgt.cols.shape                     # the reference keeps the row CDF as [h, w] (callers read [:, 0])
    for got, key in ((lgt._pdf, "pdf"), (lgt.cols, "cols"), (lgt.rows[:, 0], "rows")):
        want = g[key]
        assert got.shape == want.shape, key
        assert float((got.cpu() - want).abs().max()) <= 2e-6 * float(want.abs().max()), key
    # what the sampler relies on: CDFs end at exactly 1 and are non-decreasing
    assert float(lgt.cols[:, -1].min()) == 1.0 and float(lgt.cols[:, -1].max()) == 1.0 and float(lgt.rows[-1, 0]) == 1.0
    assert bool((lgt.cols[:, 1:] >= lgt.cols[:, :-1]).all()) and bool((lgt.rows[1:, 0] >= lgt.rows[:-1, 0]).all())
    # 256 x 256 probe of the benchmark against the torch expression of the reference (light.py:46-59)
    base = torch.rand(256, 256, 3, device=D) * 0.5 + 0.25
    l2 = light.EnvironmentLight(base)
    Y = (torch.arange(256, device=D, dtype=torch.float32) + 0.5) / 256
    pdf = base.max(-1)[0] * torch.sin(Y * np.pi)[:, None]
    pdf = pdf / pdf.sum()
    cols = torch.cumsum(pdf, 1)
    rows = torch.cumsum(cols[:, -1:].repeat(1, 256), 0)
    cols = cols / cols[:, -1:]
    rows = rows / rows[-1:, :]
    for got, want in ((l2._pdf, pdf), (l2.cols, cols), (l2.rows, rows)):
        assert float((got - want).abs().max()) <= 5e-6 * float(want.abs().max())
