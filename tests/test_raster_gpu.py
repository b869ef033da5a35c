"""GPU parity of rasterise / interpolate (C ABI) vs oracle/raster_oracle.py (PARITY UNPINNED w.r.t. nvdiffrast,
see the oracle header) + structural properties on a full-size extracted mesh."""
import math

import pytest
import torch

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu


def dev():
    return device()


def perspective(fovy=0.7854, aspect=1.0, n=0.1, f=1000.0):
    y = math.tan(fovy / 2)
    return torch.tensor([[1 / (y * aspect), 0, 0, 0], [0, 1 / -y, 0, 0], [0, 0, -(f + n) / (f - n), -(2 * f * n) / (f - n)],
                         [0, 0, -1, 0]], dtype=torch.float32)


def scene(seed, n_tri=60, B=2):
    g = torch.Generator().manual_seed(seed)
    verts = (torch.rand(n_tri * 3, 3, generator=g) - 0.5) * 1.6
    tris = torch.arange(n_tri * 3).view(-1, 3)
    # add a shared-edge fan to exercise the tie rule
    fan = torch.tensor([[0.0, 0.0, 0.2], [0.5, 0.0, 0.2], [0.5, 0.5, 0.2], [0.0, 0.5, 0.2], [-0.5, 0.5, 0.2], [-0.5, 0.0, 0.2]])
    base = verts.shape[0]
    verts = torch.cat([verts, fan], 0)
    tris = torch.cat([tris, torch.tensor([[base, base + 1, base + 2], [base, base + 2, base + 3], [base, base + 3, base + 4],
                                          [base + 5, base, base + 4]])], 0)
    mv = torch.eye(4).repeat(B, 1, 1)
    mv[:, 2, 3] = -3.0
    ang = torch.rand(B, generator=g)
    mv[:, 0, 0] = torch.cos(ang); mv[:, 0, 2] = torch.sin(ang); mv[:, 2, 0] = -torch.sin(ang); mv[:, 2, 2] = torch.cos(ang)
    mvp = perspective() @ mv
    return verts, tris, mvp


@pytest.mark.parametrize("seed,H,W", [(0, 48, 64), (1, 33, 47)])
def test_rasterize_interpolate_vs_oracle(seed, H, W):
    from gshell_b200.render import raster
    from gshell_b200.render import renderutils as ru
    from oracle import raster_oracle as ro
    from oracle import shade_oracle as so
    verts, tris, mvp = scene(seed)
    attr = torch.randn(1, verts.shape[0], 5, generator=torch.Generator().manual_seed(seed + 5))
    # ---- oracle ----
    ov = verts.clone().requires_grad_()
    oa = attr.clone().requires_grad_()
    oclip = so.xfm_points(ov[None], mvp)
    orast = ro.rasterize(oclip, tris, H, W)
    oout = ro.interpolate(oa, orast, tris)
    opos = ro.interpolate(ov[None], orast, tris)
    # ---- CUDA ----
    d = dev()
    gv = verts.clone().to(d).requires_grad_()
    ga = attr.clone().to(d).requires_grad_()
    gclip = ru.xfm_points(gv[None], mvp.to(d))
    grast, gdb = raster.rasterize(gclip, tris.to(d), (H, W))
    gout, _ = raster.interpolate(ga, grast, tris.int().to(d))
    gpos, gpos_d = raster.interpolate(gv[None], grast, tris.int().to(d), rast_db=gdb)
    assert torch.equal(grast[..., 3].cpu(), orast[..., 3].detach()), "triangle ids differ"
    cov = orast[..., 3] > 0
    assert 0.05 < cov.float().mean() < 0.95
    torch.testing.assert_close(grast[..., :3].cpu(), orast[..., :3].detach(), rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(gout.cpu(), oout.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(gpos.cpu(), opos.detach(), rtol=1e-4, atol=1e-5)
    # screen-space derivative of the interpolated position vs finite differences of the oracle along x
    # (interior pixels whose right neighbour is in the same triangle)
    same = (orast[:, :, 1:, 3] == orast[:, :, :-1, 3]) & cov[:, :, 1:]
    fd = (opos[:, :, 1:] - opos[:, :, :-1]).detach()
    an = 0.5 * (gpos_d.cpu()[:, :, 1:, 0::2] + gpos_d.cpu()[:, :, :-1, 0::2])
    err = ((fd - an).abs().max(-1).values)[same]
    assert float(err.quantile(0.99)) < 2e-3
    # ---- gradients: d/d attr, d/d verts (through interpolate weights, u/v and xfm_points) ----
    gw = torch.Generator().manual_seed(seed + 9)
    w1, w2 = torch.randn(oout.shape, generator=gw), torch.randn(opos.shape, generator=gw)
    ((oout * w1).sum() + (opos * w2).sum()).backward()
    ((gout * w1.to(d)).sum() + (gpos * w2.to(d)).sum()).backward()
    for name, a, b in (("attr", ga.grad, oa.grad), ("verts", gv.grad, ov.grad)):
        scale = b.abs().max().clamp(min=1e-6)
        assert (a.cpu() - b).abs().max() <= 2e-4 * scale, (name, float((a.cpu() - b).abs().max()), float(scale))


def test_full_size_mesh_render_properties():
    """BASELINE '256' mesh (5.1M faces) at 2 x 1024^2: ids in range, covered pixels' interpolated position lies on
    the triangle plane, barycentrics in [0,1], deterministic across runs."""
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.grids import bcc_tet_grid
    from gshell_b200.render import raster
    from gshell_b200.render import renderutils as ru
    d = dev()
    v, t = bcc_tet_grid(103)
    g = torch.Generator().manual_seed(0)
    pos = (torch.tensor(v) - 0.5).to(d)
    sdf = (torch.rand(v.shape[0], generator=g) - 0.1).to(d)
    msdf = (torch.rand(v.shape[0], generator=g) - 0.01).clamp(-1, 1).to(d)
    va, fa, _, _, _, ex = GShell_Tets(index_dtype=torch.int32, with_tangents=False)(pos, sdf, msdf, torch.tensor(t).to(d))
    _, _, mvp = scene(3)
    H = W = 1024
    clip = ru.xfm_points(va[None], mvp.to(d))
    rast, db = raster.rasterize(clip, fa, (H, W))
    rast2, _ = raster.rasterize(clip, fa, (H, W))
    assert torch.equal(rast, rast2)
    ids = rast[..., 3].long()
    assert int(ids.min()) >= 0 and int(ids.max()) <= fa.shape[0]
    cov = ids > 0
    assert float(cov.float().mean()) > 0.05
    u, vv = rast[..., 0][cov], rast[..., 1][cov]
    assert float(u.min()) >= -1e-4 and float(vv.min()) >= -1e-4 and float((u + vv).max()) <= 1 + 1e-4
    gb_pos, _ = raster.interpolate(va[None], rast, fa)
    tri = fa[(ids[cov] - 1)].long()
    p0, p1, p2 = va[tri[:, 0]], va[tri[:, 1]], va[tri[:, 2]]
    n = torch.linalg.cross(p1 - p0, p2 - p0)
    dist = ((gb_pos[cov] - p0) * n).sum(-1).abs() / n.norm(dim=-1).clamp(min=1e-12)
    assert float(dist.max()) < 1e-4
