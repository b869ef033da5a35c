"""Drop-in for the reference's `geometry/gshell_tets.py::GShell_Tets` backed by the sm_100a kernels
in csrc/mt_extract.cu (C ABI: gsb_mt_count / gsb_mt_emit / gsb_mt_backward, include/gshell_b200.h).

Same constructor and `__call__` signature / return tuple as the reference (gshell_tets.py:80-81,
245, 426-443).  Differentiable w.r.t. pos, sdf, msdf with the reference's stop-gradient structure.
CUDA tensors only: there is no CPU path in the product.
"""
import torch

from .. import _lib
from .tet_tables import tables_for

_NCOUNTS = 16


class _MarchingTets(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, sdf, msdf, tab, watertight_template=True):
        L = _lib.lib
        dev = pos.device
        stream = _lib.current_stream(dev)
        pos_c, sdf_c, msdf_c = pos.detach().contiguous(), sdf.detach().contiguous(), msdf.detach().contiguous()
        ws = tab.workspace(L.gsb_mt_workspace_bytes(tab.n_tets, tab.n_edges))
        counts, counts_host = tab.counts_buffers(_NCOUNTS)
        _lib.check(L.gsb_mt_count(_lib.ptr(pos_c), _lib.ptr(sdf_c), _lib.ptr(msdf_c), _lib.ptr(tab.tet_v), _lib.ptr(tab.tet_e),
                                  _lib.ptr(tab.edge_v), tab.n_verts, tab.n_tets, tab.n_edges, _lib.ptr(ws), ws.numel(),
                                  1 if watertight_template else 0, _lib.ptr(counts), stream), "gsb_mt_count")
        counts_host.copy_(counts, non_blocking=True)
        _lib.synchronize(dev)                                 # the ONE host sync of the extraction
        c = counts_host.tolist()
        n_wt, n_t1, n_t2 = c[0], c[1], c[2]
        g = c[3:9]
        n_aug = n_wt + 3 * n_t1 + 4 * n_t2
        n_fwt = n_t1 + 2 * n_t2
        n_faug = g[0] + 2 * g[1] + g[2] + 2 * g[3] + 3 * g[4] + 4 * g[5]
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        verts_aug = torch.empty((n_aug, 3), **f32)
        msdf_aug = torch.empty((n_aug,), **f32)
        faces_aug = torch.empty((n_faug, 3), **i32)
        verts_wt = torch.empty((n_wt, 3), **f32)
        faces_wt = torch.empty((n_fwt, 3), **i32)
        vert_edge = torch.empty((n_wt,), **i32)
        slot_a = torch.empty((n_aug - n_wt,), **i32)
        if n_wt > 0:
            _lib.check(L.gsb_mt_emit(_lib.ptr(pos_c), _lib.ptr(sdf_c), _lib.ptr(tab.tet_e), _lib.ptr(tab.edge_v),
                                     tab.n_tets, tab.n_edges, _lib.ptr(ws), _lib.ptr(counts),
                                     _lib.ptr(verts_aug), _lib.ptr(msdf_aug), _lib.ptr(faces_aug),
                                     _lib.ptr(verts_wt), _lib.ptr(faces_wt), _lib.ptr(vert_edge),
                                     _lib.ptr(slot_a), stream), "gsb_mt_emit")
        ctx.tab = tab
        ctx.sizes = (n_wt, n_t1, n_t2)
        ctx.save_for_backward(pos_c, sdf_c, msdf_c, verts_wt, msdf_aug, vert_edge, slot_a)
        ctx.mark_non_differentiable(faces_aug, faces_wt, slot_a)
        return verts_aug, msdf_aug, verts_wt, faces_aug, faces_wt, slot_a

    @staticmethod
    def backward(ctx, g_verts_aug, g_msdf_aug, g_verts_wt, _gfa, _gfw, _gsa):
        L = _lib.lib
        pos, sdf, msdf, verts_wt, msdf_aug, vert_edge, slot_a = ctx.saved_tensors
        tab = ctx.tab
        n_wt, n_t1, n_t2 = ctx.sizes
        dev = pos.device
        g_pos = torch.zeros_like(pos)
        g_sdf = torch.zeros_like(sdf)
        g_msdf = torch.zeros_like(msdf)
        if n_wt > 0:
            def prep(g):
                return None if g is None else g.contiguous().float()
            ga, gm, gw = prep(g_verts_aug), prep(g_msdf_aug), prep(g_verts_wt)
            scratch = torch.empty((n_wt, 5), dtype=torch.float32, device=dev)
            _lib.check(L.gsb_mt_backward(_lib.ptr(pos), _lib.ptr(sdf), _lib.ptr(msdf), _lib.ptr(tab.edge_v),
                                         _lib.ptr(verts_wt), _lib.ptr(msdf_aug), _lib.ptr(vert_edge),
                                         _lib.ptr(slot_a), n_wt, n_t1, n_t2,
                                         _lib.ptr(ga), _lib.ptr(gm), _lib.ptr(gw), _lib.ptr(scratch),
                                         _lib.ptr(g_pos), _lib.ptr(g_sdf), _lib.ptr(g_msdf),
                                         _lib.current_stream(dev)), "gsb_mt_backward")
        return g_pos, g_sdf, g_msdf, None, None


def _n_tri_polys(n_boundary, n_faces_wt):
    """T1 from the sizes: n_boundary = 3 T1 + 4 T2, n_faces_wt = T1 + 2 T2  =>  T1 = 2 n_faces_wt... solved exactly."""
    # 3 T1 + 4 T2 = nb ; T1 + 2 T2 = nf  =>  T1 = nb - 2 nf
    return n_boundary - 2 * n_faces_wt


class GShell_Tets:
    """`GShell_Tets()(pos_nx3, sdf_n, msdf_n, tet_fx4) -> (verts_aug, faces_aug, None, None, v_tng_aug, extra)`.

    Options beyond the reference (keyword-only, defaults reproduce the reference):
      index_dtype   dtype of the returned face tensors.  The reference returns int64 and every consumer
                    narrows to int32 (`.int()` at render.py:240, gshell_tets_geometry.py:211); the
                    kernels emit int32, so `torch.int32` skips a widening pass.
      with_tangents compute v_tng_aug (dead on the training path; the geometry module turns it
                    off, exactly as the reference's getMesh discards it).
    """

    def __init__(self, index_dtype=torch.int64, with_tangents=True):
        self.index_dtype = index_dtype
        self.with_tangents = with_tangents

    def __call__(self, pos_nx3, sdf_n, msdf_n, tet_fx4, output_watertight_template=True):
        _lib.require_cuda(pos_nx3, "gshell_b200.GShell_Tets")
        tab = tables_for(tet_fx4, pos_nx3.shape[0])
        sdf = sdf_n.float().reshape(-1)
        msdf = msdf_n.float().reshape(-1)
        # output_watertight_template=False (reference :260-263, no caller in the reference): tets without a positive mSDF corner are
        # dropped BEFORE the vertex numbering (fewer vertices / rows, other ids) and `extra` keeps only its mSDF entries (:435-441)
        verts_aug, msdf_aug, verts_wt, faces_aug, faces_wt, slot_a = _MarchingTets.apply(pos_nx3.float(), sdf, msdf, tab,
                                                                                         bool(output_watertight_template))
        n_wt = verts_wt.shape[0]
        if self.index_dtype != torch.int32:
            faces_aug, faces_wt = faces_aug.to(self.index_dtype), faces_wt.to(self.index_dtype)
        v_tng = v_tng_aug = None
        if self.with_tangents:
            from .tangents import tangent_frame_aug
            v_tng, v_tng_aug = tangent_frame_aug(verts_wt, faces_wt, msdf_aug[:n_wt], slot_a, tab.n_tets,
                                                 _n_tri_polys(verts_aug.shape[0] - n_wt, faces_wt.shape[0]),
                                                 sdf=sdf, msdf=msdf, edge_v=tab.edge_v)
        extra = {
            "n_verts_watertight": n_wt,
            "vertices_watertight": verts_wt,
            "faces_watertight": faces_wt,
            "v_tng_watertight": v_tng,
            "msdf": msdf_aug,
            "msdf_watertight": msdf_aug[:n_wt],
            "msdf_boundary": msdf_aug[n_wt:],
        }
        if not output_watertight_template:
            extra = {k: extra[k] for k in ("msdf", "msdf_watertight", "msdf_boundary")}
        return verts_aug, faces_aug, None, None, v_tng_aug, extra

    # ---- generative decode path ------------------------------------------------------------------------------------
    @torch.no_grad()
    def marching_from_auggrid(self, pos_nx3, sdf_n, tet_fx4, sorted_tet_edges_fx6x2, coeff_sdf_interp, verts_discretized,
                              midpoint_msdf_sign_n, occgrid):
        """Reference gshell_tets.py:446-629 (decode of generated augmented grids; no gradients there either).  Returns the
        reference's 9-tuple (verts_aug, faces_aug, None, None, v_tng_aug, verts, valid_tet_gidx, msdf_vert_aug, msdf_vert).
        Five kernels (csrc/auggrid.cu) around two int32 prefix sums: the per-call `unique(dim=0)` over the valid tets' edges
        (:467) is the static sorted edge table of the grid plus a scan (same numbering, see tet_tables.py); two host reads size
        the outputs."""
        _lib.require_cuda(pos_nx3, "gshell_b200.GShell_Tets")
        return self._marching_from_auggrid(pos_nx3, sdf_n, tet_fx4, sorted_tet_edges_fx6x2, coeff_sdf_interp, verts_discretized,
                                           midpoint_msdf_sign_n, occgrid)

    def _marching_from_auggrid(self, pos, sdf_n, tet_fx4, sorted_tet_edges, coeff_grid, verts_disc, msdf_sign_grid, occgrid):
        import ctypes
        from .mt_luts import luts_i32
        L = _lib.lib
        dev = pos.device
        stream = _lib.current_stream(dev)
        tab = tables_for(tet_fx4, pos.shape[0])
        if not getattr(tab, "_auggrid_edges_checked", False):
            # the reference numbers vertices through the caller's per-tet edge list; the static table is equivalent only if that
            # list is what the reference's own tables assume: the 6 base edges of every tet with sorted endpoints
            if not torch.equal(sorted_tet_edges.reshape(-1, 6, 2).long(), tab.edge_v.long()[tab.tet_e.long()]):
                raise ValueError("sorted_tet_edges_fx6x2 must hold, per tet, the edges (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) of "
                                 "tet_fx4 with sorted endpoints")
            tab._auggrid_edges_checked = True

        def f32(t):
            return t.detach().float().contiguous()
        p, sdf, disc = f32(pos), f32(sdf_n).reshape(-1), f32(verts_disc)
        coeff, msign, occ = f32(coeff_grid), f32(msdf_sign_grid), f32(occgrid)
        if coeff.dim() != 3 or msign.shape != coeff.shape or occ.dim() != 3:
            raise ValueError("coeff_sdf_interp / midpoint_msdf_sign_n must be equally sized 3-D grids, occgrid a 3-D grid")
        n_edges, n_tets = tab.n_edges, tab.n_tets
        i32 = dict(dtype=torch.int32, device=dev)
        fl = dict(dtype=torch.float32, device=dev)
        # ---- vertices: the crossing edges of the static sorted edge table, in table order (= the reference's `unique`, :467) ------
        flags = torch.empty((n_edges,), **i32)
        _lib.check(L.gsb_auggrid_edge_flags(_lib.ptr(sdf), _lib.ptr(tab.edge_v), n_edges, _lib.ptr(flags), stream), "gsb_auggrid_edge_flags")
        flags_incl = torch.cumsum(flags, 0, dtype=torch.int32)
        n_wt = int(flags_incl[-1]) if n_edges > 0 else 0                         # host read: sizes the vertex arrays
        verts, cano, m_vert = torch.empty((n_wt, 3), **fl), torch.empty((n_wt, 3), **fl), torch.empty((n_wt,), **fl)
        _lib.check(L.gsb_auggrid_vertices(_lib.ptr(p), _lib.ptr(disc), _lib.ptr(tab.edge_v), _lib.ptr(flags), _lib.ptr(flags_incl), n_edges,
                                          _lib.ptr(coeff), _lib.ptr(msign), *coeff.shape, _lib.ptr(verts), _lib.ptr(cano), _lib.ptr(m_vert),
                                          stream), "gsb_auggrid_vertices")
        # ---- polygons and cut groups per tet, then their output rows ----------------------------------------------------------------
        lut = luts_i32(dev)
        lut_table = (ctypes.c_void_p * 7)(*[t.data_ptr() for t in lut])
        rows = torch.empty((8, n_tets), **i32)
        _lib.check(L.gsb_auggrid_classify(_lib.ptr(sdf), _lib.ptr(tab.tet_v), _lib.ptr(tab.tet_e), _lib.ptr(flags), _lib.ptr(flags_incl),
                                          _lib.ptr(m_vert), n_tets, lut_table, _lib.ptr(rows), stream), "gsb_auggrid_classify")
        rows_incl = torch.cumsum(rows, 1, dtype=torch.int32)
        totals = [int(x) for x in rows_incl[:, -1].tolist()] if n_tets > 0 else [0] * 8      # host read: sizes the outputs
        n_one, n_two = totals[0], totals[1]
        nb = 3 * n_one + 4 * n_two
        n_faug = sum(c * k for c, k in zip(totals[2:], (1, 2, 1, 2, 3, 4)))
        faces_wt, tet_ids = torch.empty((n_one + 2 * n_two, 3), **i32), torch.empty((n_one + n_two,), **i32)
        b_pos, b_ab, b_w = torch.empty((nb, 3), **fl), torch.empty((nb, 2), **i32), torch.empty((nb, 2), **fl)
        faces_aug = torch.empty((n_faug, 3), **i32)
        _lib.check(L.gsb_auggrid_emit(_lib.ptr(sdf), _lib.ptr(tab.tet_v), _lib.ptr(tab.tet_e), _lib.ptr(flags), _lib.ptr(flags_incl),
                                      _lib.ptr(m_vert), _lib.ptr(verts), _lib.ptr(cano), _lib.ptr(rows_incl), n_tets, lut_table, _lib.ptr(occ),
                                      *occ.shape, n_wt, (ctypes.c_int64 * 8)(*totals), _lib.ptr(faces_wt), _lib.ptr(tet_ids), _lib.ptr(b_pos),
                                      _lib.ptr(b_ab), _lib.ptr(b_w), _lib.ptr(faces_aug), stream), "gsb_auggrid_emit")
        verts_aug = torch.cat([verts, b_pos], 0)
        v_tng_aug = None
        if self.with_tangents:
            from .tangents import tangent_frame_wt
            v_tng = tangent_frame_wt(verts, faces_wt, n_tets)
            b_tng = torch.empty((nb, 3), **fl)
            _lib.check(L.gsb_auggrid_boundary_attr(_lib.ptr(v_tng.contiguous()), _lib.ptr(b_ab), _lib.ptr(b_w), nb, _lib.ptr(b_tng), stream),
                       "gsb_auggrid_boundary_attr")
            v_tng_aug = torch.cat([v_tng, b_tng], 0)
        m_aug = torch.cat([m_vert, torch.zeros((nb,), **fl)])
        return verts_aug, faces_aug.to(self.index_dtype), None, None, v_tng_aug, verts, tet_ids.long(), m_aug, m_vert
