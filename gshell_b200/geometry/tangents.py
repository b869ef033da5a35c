"""Tangent frame of the watertight mesh and its extension to the boundary vertices (reference gshell_tets.py:40-78,
210-239, 318-319, 337-338, 375-380), on the kernels of csrc/tangents.cu (C ABI: gsb_tangents_fwd / gsb_tangents_bwd) and the
vertex-normal kernels -- 5 + 6 launches instead of ~40 ATen kernels and a 4 * ceil(sqrt(T))^2-row uv table (415 MB at the
"256" grid).  This output is dead on the training path (getMesh drops it; render_layer synthesises a random tangent).

Reference quirk kept on purpose: `compute_tangents(verts, uvs_pre, v_nrm, faces, faces, faces)` indexes the uv atlas with
the VERTEX ids of each face (not with the per-face uv indices it just built), so the "uv" of vertex v is atlas row v."""
import math

import torch

from .. import _lib
from ..render.mesh import vertex_normals


def _atlas(n_tets, device):
    """The atlas map_uv builds (:210-225) is 4 corners per cell of an N x N grid, N = ceil(sqrt(n_tets)): the kernels only need
    the N cell origins (the reference's own linspace, so that the values are its values) and the corner offset."""
    n = int(math.ceil(math.sqrt((2 * n_tets + 1) // 2)))
    return torch.linspace(0, 1 - (1 / n), n, dtype=torch.float32, device=device), n, 0.9 / n


class _TangentFrame(torch.autograd.Function):
    """(verts_wt, normals_wt, msdf_wt | None) -> tangents [Vw + nb, 3]; differentiable w.r.t. all three."""

    @staticmethod
    def forward(ctx, verts, nrm, msdf_wt, faces, slot_a, n_tets, n_tri_polys):
        dev = verts.device
        v, n = verts.detach().float().contiguous(), nrm.detach().float().contiguous()
        f = faces.int().contiguous()
        m = None if msdf_wt is None else msdf_wt.detach().float().contiguous()
        s = None if slot_a is None else slot_a.int().contiguous()
        n_wt, nb = v.shape[0], (0 if s is None else s.shape[0])
        lin, an, pad = _atlas(n_tets, dev)
        acc = torch.empty((n_wt, 4), dtype=torch.float32, device=dev)
        tng = torch.empty((n_wt + nb, 3), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib.gsb_tangents_fwd(_lib.ptr(v), _lib.ptr(f), _lib.ptr(n), _lib.ptr(m), _lib.ptr(s), _lib.ptr(lin), n_wt,
                                             f.shape[0], int(n_tri_polys), nb, an, pad, _lib.ptr(acc), _lib.ptr(tng),
                                             _lib.current_stream(dev)), "gsb_tangents_fwd")
        ctx.save_for_backward(f, n, m, s, lin, acc, tng)
        ctx.meta = (int(n_tri_polys), an, pad)
        return tng

    @staticmethod
    def backward(ctx, g_tng):
        f, n, m, s, lin, acc, tng = ctx.saved_tensors
        n_tri_polys, an, pad = ctx.meta
        dev = n.device
        n_wt, nb = n.shape[0], (0 if s is None else s.shape[0])
        g = g_tng.float().contiguous()
        g_t, g_v, g_n = (torch.empty((n_wt, 3), dtype=torch.float32, device=dev) for _ in range(3))
        g_m = torch.empty((n_wt,), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib.gsb_tangents_bwd(_lib.ptr(f), _lib.ptr(n), _lib.ptr(m), _lib.ptr(s), _lib.ptr(lin), n_wt, f.shape[0],
                                             n_tri_polys, nb, an, pad, _lib.ptr(acc), _lib.ptr(tng), _lib.ptr(g), _lib.ptr(g_t),
                                             _lib.ptr(g_v), _lib.ptr(g_n), _lib.ptr(g_m), _lib.current_stream(dev)),
                   "gsb_tangents_bwd")
        return g_v, g_n, (g_m if m is not None else None), None, None, None, None


def tangent_frame_wt(verts_wt, faces_wt, n_tets):
    """Per-vertex tangents of the watertight mesh (compute_tangents :40-78 on the map_uv atlas) -> [Vw,3]."""
    if verts_wt.shape[0] == 0:
        return torch.zeros((0, 3), dtype=torch.float32, device=verts_wt.device)
    return _TangentFrame.apply(verts_wt, vertex_normals(verts_wt, faces_wt), None, faces_wt, None, n_tets, 0)


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
    msdf_wt = _msdf_with_sdf_gradient(msdf_wt, sdf, msdf, edge_v)
    # boundary vertices: same mSDF zero-crossing weights as the positions (:345-365, :375-380), inside the kernel
    tng = _TangentFrame.apply(verts_wt, vertex_normals(verts_wt, faces_wt), msdf_wt, faces_wt, slot_a, n_tets, n_tri_polys)
    return tng[:n_wt], tng
