"""Loss terms of one training iteration on the CUDA reductions of csrc/tick_ops.cu (C ABI: gsb_sdf_reg_*, gsb_msdf_reg_*,
gsb_image_terms_*), each with its hand-written adjoint behind a torch.autograd.Function.

  sdf_reg_loss            reference geometry/gshell_tets_geometry.py:33-39 (compute_sdf_reg_loss)
  msdf_reg_loss           reference geometry/gshell_tets_geometry.py:325-356 (open / close mSDF Huber terms)
  image_terms             reference geometry/gshell_tets_geometry.py:283-290 (alpha MSE, mSDF-image L1 terms) and
                          render/regularizer.py:21-52 (chroma_loss, shading_loss, material_smoothness_grad) in one pass
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib

T_ALPHA, T_MSDF, T_CHROMA, T_SHADING, T_SMOOTH = 1, 2, 4, 8, 16


def _cuda_f32(t):
    _lib.require_cuda(t, "gshell_b200.losses")
    return t.detach().float().contiguous()


class _SdfReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, edges):
        s = _cuda_f32(sdf).reshape(-1)
        acc = torch.empty(2, dtype=torch.float64, device=s.device)
        _lib.check(_lib.lib.gsb_sdf_reg_fwd(_lib.ptr(s), _lib.ptr(edges), edges.shape[0], _lib.ptr(acc), _lib.current_stream(s.device)),
                   "gsb_sdf_reg_fwd")
        ctx.save_for_backward(s, edges, acc)
        ctx.shape = sdf.shape
        return (acc[0] / acc[1]).float()          # mean over the sign-changing edges (nan when there is none, like the reference)

    @staticmethod
    def backward(ctx, g):
        s, edges, acc = ctx.saved_tensors
        gs = torch.zeros_like(s)
        _lib.check(_lib.lib.gsb_sdf_reg_bwd(_lib.ptr(s), _lib.ptr(edges), edges.shape[0], _lib.ptr(acc), _lib.ptr(g.float().contiguous()),
                                            1.0, _lib.ptr(gs), _lib.current_stream(s.device)), "gsb_sdf_reg_bwd")
        return gs.view(ctx.shape), None


def sdf_reg_loss(sdf, edges):
    """BCE between the SDF values at the two ends of every sign-changing grid edge; `edges` int32 [E,2] (static table)."""
    assert edges.dtype == torch.int32 and edges.is_contiguous()
    return _SdfReg.apply(sdf, edges)


def visible_boundary_mask(tris, visible_triangles, n_verts_watertight, n_boundary):
    """uint8 [n_boundary]: boundary vertices (ids >= n_verts_watertight) referenced by any visible triangle."""
    bmask = torch.zeros(n_boundary, dtype=torch.uint8, device=tris.device)
    vis = visible_triangles.long().contiguous()
    if vis.numel():
        _lib.check(_lib.lib.gsb_mark_visible_boundary(_lib.ptr(tris.int().contiguous()), _lib.ptr(vis), vis.numel(), int(n_verts_watertight),
                                                      _lib.ptr(bmask), _lib.current_stream(tris.device)), "gsb_mark_visible_boundary")
    return bmask


class _MsdfReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, msdf_all, msdf_boundary, bmask, eps, w_open, w_close):
        ma = _cuda_f32(msdf_all).reshape(-1) if (msdf_all is not None and w_open != 0) else None
        mb = _cuda_f32(msdf_boundary).reshape(-1) if (msdf_boundary is not None and bmask is not None and w_close != 0) else None
        dev = (ma if ma is not None else mb).device
        acc = torch.empty(2, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib.gsb_msdf_reg_fwd(_lib.ptr(ma), 0 if ma is None else ma.numel(), _lib.ptr(mb), _lib.ptr(bmask),
                                             0 if mb is None else mb.numel(), float(eps), _lib.ptr(acc), _lib.current_stream(dev)),
                   "gsb_msdf_reg_fwd")
        ctx.saved = (ma, mb, bmask)
        ctx.meta = (float(eps), float(w_open), float(w_close), None if msdf_all is None else msdf_all.shape,
                    None if msdf_boundary is None else msdf_boundary.shape)
        return (acc[0] * w_open + acc[1] * w_close).float()

    @staticmethod
    def backward(ctx, g):
        ma, mb, bmask = ctx.saved
        eps, w_open, w_close, shape_a, shape_b = ctx.meta
        dev = (ma if ma is not None else mb).device
        ga = torch.empty_like(ma) if (ma is not None and ctx.needs_input_grad[0]) else None
        gb = torch.empty_like(mb) if (mb is not None and ctx.needs_input_grad[1]) else None
        _lib.check(_lib.lib.gsb_msdf_reg_bwd(_lib.ptr(ma), 0 if ma is None else ma.numel(), _lib.ptr(mb), _lib.ptr(bmask),
                                             0 if mb is None else mb.numel(), eps, _lib.ptr(g.float().contiguous()), w_open, w_close,
                                             _lib.ptr(ga), _lib.ptr(gb), _lib.current_stream(dev)), "gsb_msdf_reg_bwd")
        return (None if ga is None else ga.view(shape_a), None if gb is None else gb.view(shape_b), None, None, None, None)


def msdf_reg_loss(msdf_all, msdf_boundary, bmask, w_open, w_close, eps=1e-3):
    """w_open * sum huber(max(msdf, -eps), -eps) over all mesh vertices + w_close * sum huber(min(msdf_b, eps), eps) over the
    boundary vertices selected by `bmask` (uint8)."""
    return _MsdfReg.apply(msdf_all, msdf_boundary, bmask, eps, w_open, w_close)


def _as4(t):
    """[B,H,W,3|4] -> contiguous [B,H,W,4] (zero alpha appended to 3-channel inputs)."""
    if t is None:
        return None
    if t.shape[-1] == 3:
        t = torch.nn.functional.pad(t, (0, 1))
    assert t.shape[-1] == 4
    return t.float().contiguous()


class _ImageTerms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shaded, msdf_img, kd, kd_grad, ks_grad, nrm_grad, diffuse, specular, ref, terms, lambdas):
        tens = [None if t is None else t.detach() for t in (shaded, msdf_img, kd, kd_grad, ks_grad, nrm_grad, diffuse, specular, ref)]
        for i, t in enumerate(tens):
            if t is not None:
                _lib.require_cuda(t, "gshell_b200.losses")
                assert t.is_contiguous() and t.dtype == torch.float32 and (i == 1 or t.shape[-1] == 4)
        ref_t = tens[8]
        dev = ref_t.device
        n_pix = ref_t.numel() // 4
        msdf_ch = 0 if tens[1] is None else tens[1].shape[-1]
        lam = (ctypes.c_float * 6)(*[float(x) for x in lambdas])
        acc = torch.empty(int(_lib.lib.gsb_image_terms_accumulators()), dtype=torch.float64, device=dev)
        stream = _lib.current_stream(dev)
        ptrs = [_lib.ptr(t) for t in tens]
        _lib.check(_lib.lib.gsb_image_terms_reduce(*ptrs, n_pix, msdf_ch, int(terms), lam, _lib.ptr(acc), stream), "gsb_image_terms_reduce")
        mean_scale = 1.0
        if (terms & T_SHADING) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # the specular / diffuse means of shading_loss are over the whole batch (views shard across ranks)
            dist.all_reduce(acc[5:7])
            mean_scale = 1.0 / dist.get_world_size()
        out = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib.gsb_image_terms_finish(ptrs[1], ptrs[6], ptrs[7], n_pix, msdf_ch, int(terms), lam, _lib.ptr(acc), mean_scale,
                                                   _lib.ptr(out), stream), "gsb_image_terms_finish")
        ctx.tens, ctx.acc, ctx.meta = tens, acc, (n_pix, msdf_ch, int(terms), lam, mean_scale)
        return out

    @staticmethod
    def backward(ctx, g_out):
        tens, acc = ctx.tens, ctx.acc
        n_pix, msdf_ch, terms, lam, mean_scale = ctx.meta
        dev = acc.device
        grads = [torch.empty_like(t) if (t is not None and ctx.needs_input_grad[i]) else None for i, t in enumerate(tens[:8])]
        _lib.check(_lib.lib.gsb_image_terms_bwd(*[_lib.ptr(t) for t in tens], n_pix, msdf_ch, terms, lam, _lib.ptr(acc), mean_scale,
                                                _lib.ptr(g_out.float().contiguous()), *[_lib.ptr(g) for g in grads],
                                                _lib.current_stream(dev)), "gsb_image_terms_bwd")
        return (*grads, None, None, None)


def image_terms(ref, terms, lambdas, shaded=None, msdf_img=None, kd=None, kd_grad=None, ks_grad=None, nrm_grad=None, diffuse=None,
                specular=None):
    """-> float32 [2]: (image part, regulariser part).  Buffers are the composited [B,H,W,4] images of render_mesh
    (msdf_img [B,H,W,C]); lambdas = (chroma, diffuse, specular, kd, ks, nrm)."""
    m = None if msdf_img is None else msdf_img.float().contiguous()
    return _ImageTerms.apply(_as4(shaded), m, _as4(kd), _as4(kd_grad), _as4(ks_grad), _as4(nrm_grad), _as4(diffuse), _as4(specular),
                             _as4(ref), terms, lambdas)
