"""Tangent frame of the watertight mesh and its extension to the boundary vertices (reference gshell_tets.py:40-78,
210-239, 318-319, 337-338, 375-380).  This output is dead on the training path (getMesh drops it; render_layer
synthesises a random tangent), so it is composed from torch ops on the GPU plus the vertex-normal kernel; autograd
provides its gradients.

Reference quirk kept on purpose: `compute_tangents(verts, uvs_pre, v_nrm, faces, faces, faces)` indexes the uv atlas with
the VERTEX ids of each face (not with the per-face uv indices it just built), so the "uv" of vertex v is atlas row v."""
import math

import torch

from ..render.mesh import vertex_normals


def _dot(a, b):
    return (a * b).sum(-1, keepdim=True)


def _unit(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(_dot(x, x), min=eps))


def _atlas_uv(idx, n_tets, device):
    """Row `idx` of the atlas built by map_uv (:210-225): 4 corners per cell of an N x N grid, N = ceil(sqrt(n_tets))."""
    n = int(math.ceil(math.sqrt((2 * n_tets + 1) // 2)))
    lin = torch.linspace(0, 1 - (1 / n), n, dtype=torch.float32, device=device)
    pad = 0.9 / n
    cell, corner = torch.div(idx, 4, rounding_mode="floor"), idx % 4
    x, y = lin[cell % n], lin[torch.div(cell, n, rounding_mode="floor")]
    x = torch.where((corner == 1) | (corner == 2), x + pad, x)
    y = torch.where(corner >= 2, y + pad, y)
    return torch.stack([x, y], -1)


def tangent_frame_wt(verts_wt, faces_wt, n_tets):
    """Per-vertex tangents of the watertight mesh (compute_tangents :40-78 on the map_uv atlas) -> [Vw,3]."""
    dev = verts_wt.device
    f = faces_wt.long()
    nrm = vertex_normals(verts_wt, faces_wt)
    p = [verts_wt[f[:, i]] for i in range(3)]
    t = [_atlas_uv(f[:, i], n_tets, dev) for i in range(3)]
    du1, du2 = t[1] - t[0], t[2] - t[0]
    dp1, dp2 = p[1] - p[0], p[2] - p[0]
    nom = dp1 * du2[:, 1:2] - dp2 * du1[:, 1:2]
    den = du1[:, 0:1] * du2[:, 1:2] - du1[:, 1:2] * du2[:, 0:1]
    tang = nom / torch.where(den > 0, torch.clamp(den, min=1e-6), torch.clamp(den, max=-1e-6))
    acc = torch.zeros_like(nrm)
    cnt = torch.zeros_like(nrm)
    for i in range(3):
        acc = acc.index_add(0, f[:, i], tang)
        cnt = cnt.index_add(0, f[:, i], torch.ones_like(tang))
    tng = _unit(acc / cnt)
    return _unit(tng - _dot(tng, nrm) * nrm)


def _msdf_with_sdf_gradient(msdf_wt, sdf, msdf, edge_v):
    """The values of `msdf_wt` with the GRADIENT of the reference's `msdf_vert` (:287-288).

    The extraction hands out the interpolated mSDF of the crossing vertices in its stop-gradient form (`extra['msdf']`,
    reference `msdf_vert_stopvgd` :289: no gradient through the SDF interpolation weights).  The boundary weights of the
    tangents (and of the positions, inside the extraction's own adjoint) are built from `msdf_vert` itself, whose gradient also
    reaches the SDF.  That path is rebuilt here from the grid values: vertex i sits on the i-th sign-changing edge of the static
    sorted edge table (the order the reference's `unique` yields), m_i = (msdf_lo * (-sdf_hi) + msdf_hi * sdf_lo) / den."""
    if sdf is None or msdf is None or edge_v is None:
        return msdf_wt
    ev = edge_v.long()
    inside = sdf.detach() > 0
    edge = ev[inside[ev[:, 0]] != inside[ev[:, 1]]]
    if edge.shape[0] != msdf_wt.shape[0]:
        # output_watertight_template=False numbers only the crossing edges of the surviving tets (:260-263); its callers get
        # the stop-gradient form
        return msdf_wt
    s_lo, s_hi = sdf[edge[:, 0]], -sdf[edge[:, 1]]
    den = s_lo + s_hi
    den = torch.sign(den) * (den.abs() + 1e-12)
    den = torch.where(den == 0, torch.full_like(den, 1e-12), den)
    m = msdf[edge[:, 0]] * (s_hi / den) + msdf[edge[:, 1]] * (s_lo / den)
    return msdf_wt.detach() + (m - m.detach())        # value: bit-identical to the kernel's; gradient: the full path


def tangent_frame_aug(verts_wt, faces_wt, msdf_wt, slot_a, n_tets, n_tri_polys, sdf=None, msdf=None, edge_v=None):
    """-> (v_tng [Vw,3], v_tng_aug [Va,3]).  `sdf`, `msdf` (grid values, differentiable) and `edge_v` (static sorted edge
    table) let the boundary weights carry the reference's gradient to the SDF; without them the weights see the mSDF only."""
    dev = verts_wt.device
    n_wt = verts_wt.shape[0]
    if n_wt == 0:
        z = torch.zeros((0, 3), device=dev)
        return z, z
    v_tng = tangent_frame_wt(verts_wt, faces_wt, n_tets)
    msdf_wt = _msdf_with_sdf_gradient(msdf_wt, sdf, msdf, edge_v)
    # boundary vertices: same mSDF zero-crossing weights as the positions (:345-365, :375-380)
    a = (slot_a & 0x7FFFFFFF).long()
    n3 = 3 * n_tri_polys
    b = torch.cat([a[:n3].view(-1, 3).roll(-1, 1).reshape(-1), a[n3:].view(-1, 4).roll(-1, 1).reshape(-1)])
    ma, mb = msdf_wt[a], msdf_wt[b]
    ok = ((torch.sign(ma) + torch.sign(mb)).abs() != 2) & ((ma - mb).abs() > 1e-12)
    den_m = torch.where(ok, ma - mb, torch.ones_like(ma))
    w0 = torch.where(ok, -mb / den_m, torch.zeros_like(ma)).unsqueeze(-1)
    w1 = torch.where(ok, ma / den_m, torch.zeros_like(ma)).unsqueeze(-1)
    return v_tng, torch.cat([v_tng, v_tng[a] * w0 + v_tng[b] * w1], 0)
