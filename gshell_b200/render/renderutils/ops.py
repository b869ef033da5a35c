"""Autograd wrappers over the sm_100a G-buffer kernels (csrc/gbuffer_ops.cu) with the signatures of the
reference's render/renderutils/ops.py (xfm_points :518, prepare_shading_normal :197, image_loss :479).
CUDA tensors only; `use_python=True` (the reference's validation switch) is refused here -- the PyTorch
restatements live in oracle/shade_oracle.py, outside the product."""
import ctypes

import torch

from ... import _lib

_LOSS = {"l1": 0, "mse": 1, "relmse": 2, "smape": 3}
_TONEMAP = {"none": 0, "log_srgb": 1}


def _no_python(flag):
    if flag:
        raise NotImplementedError("gshell_b200 has no PyTorch fallback path; see oracle/shade_oracle.py for the checker")


def _need_cuda(t, what):
    _lib.require_cuda(t, what)


# ------------------------------------------------------------------------------------------------
class _XfmPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, matrix):
        p = points.detach().float().contiguous()
        m = matrix.detach().float().contiguous()
        B, N = m.shape[0], p.shape[1]
        batched = 1 if p.shape[0] != 1 else 0
        if batched and p.shape[0] != B:
            raise RuntimeError("xfm_points: points batch must be 1 or match the matrix batch")
        out = torch.empty((B, N, 4), dtype=torch.float32, device=p.device)
        _lib.check(_lib.lib.gsb_xfm_points_fwd(_lib.ptr(p), _lib.ptr(m), B, N, batched, _lib.ptr(out),
                                               _lib.current_stream(p.device)), "gsb_xfm_points_fwd")
        ctx.save_for_backward(m)
        ctx.meta = (B, N, batched, points.shape)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (m,) = ctx.saved_tensors
        B, N, batched, shape = ctx.meta
        g = g_out.float().contiguous()
        g_pts = torch.empty(shape, dtype=torch.float32, device=g.device)
        _lib.check(_lib.lib.gsb_xfm_points_bwd(_lib.ptr(m), _lib.ptr(g), B, N, batched, _lib.ptr(g_pts),
                                               _lib.current_stream(g.device)), "gsb_xfm_points_bwd")
        return g_pts, None


def xfm_points(points, matrix, use_python=False):
    """[1|B,N,3] x [B,4,4] -> homogeneous [B,N,4] (reference ops.py:518-533)."""
    _no_python(use_python)
    _need_cuda(points, "xfm_points")
    return _XfmPoints.apply(points, matrix)


# ------------------------------------------------------------------------------------------------
def _expand3(t, B, H, W):
    """View of t broadcast to [B,H,W,3] with unit channel stride; returns (tensor, strides b/y/x)."""
    e = t.float().expand(B, H, W, 3)
    if e.stride(3) != 1 and e.shape[3] != 1:
        e = e.contiguous()
    return e, (e.stride(0), e.stride(1), e.stride(2))


class _ShadingNormal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided, opengl):
        ins = [pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm]
        B = max(t.shape[0] for t in ins)
        H = max(t.shape[1] for t in ins)
        W = max(t.shape[2] for t in ins)
        views, strides = [], []
        for t in ins:
            e, s = _expand3(t.detach(), B, H, W)
            views.append(e)
            strides += list(s)
        out = torch.empty((B, H, W, 3), dtype=torch.float32, device=pos.device)
        ptrs = (ctypes.c_void_p * 6)(*[v.data_ptr() for v in views])
        st = (ctypes.c_int64 * 18)(*strides)
        _lib.check(_lib.lib.gsb_shading_normal_fwd(ptrs, st, B, H, W, int(two_sided), int(opengl), _lib.ptr(out),
                                                   _lib.current_stream(pos.device)), "gsb_shading_normal_fwd")
        ctx.save_for_backward(*views)
        ctx.meta = (B, H, W, strides, int(two_sided), int(opengl), [t.shape for t in ins])
        return out

    @staticmethod
    def backward(ctx, g_out):
        views = ctx.saved_tensors
        B, H, W, strides, two_sided, opengl, shapes = ctx.meta
        g = g_out.float().contiguous()
        need = ctx.needs_input_grad[:6]
        bufs = [torch.empty((B, H, W, 3), dtype=torch.float32, device=g.device) if n else None for n in need]
        ptrs = (ctypes.c_void_p * 6)(*[v.data_ptr() for v in views])
        st = (ctypes.c_int64 * 18)(*strides)
        gptrs = (ctypes.c_void_p * 6)(*[None if b is None else b.data_ptr() for b in bufs])
        _lib.check(_lib.lib.gsb_shading_normal_bwd(ptrs, st, B, H, W, two_sided, opengl, _lib.ptr(g), gptrs,
                                                   _lib.current_stream(g.device)), "gsb_shading_normal_bwd")
        # broadcast inputs: reduce the full-resolution gradient to the input's shape (as autograd does for
        # the reference's full-res outputs, ops.py:189-191)
        outs = [None if b is None else b.sum_to_size(shp) for b, shp in zip(bufs, shapes)]
        return (*outs, None, None)


def prepare_shading_normal(pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading=True,
                           opengl=True, use_python=False):
    """Reference ops.py:197-236: tangent-frame perturbation, two-sided flip, bent normal."""
    _no_python(use_python)
    _need_cuda(pos, "prepare_shading_normal")
    if perturbed_nrm is None:
        perturbed_nrm = torch.tensor([0, 0, 1], dtype=torch.float32, device=pos.device)[None, None, None, :]
    return _ShadingNormal.apply(pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm, two_sided_shading, opengl)


# ------------------------------------------------------------------------------------------------
class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, target, loss, tonemapper):
        shape = torch.broadcast_shapes(img.shape, target.shape)
        a = img.detach().float().expand(shape).contiguous()
        b = target.detach().float().expand(shape).contiguous()
        n = a.numel()
        npix = n // 3
        nb = _lib.lib.gsb_image_loss_partials(n)
        partial = torch.empty((nb,), dtype=torch.float32, device=a.device)
        _lib.check(_lib.lib.gsb_image_loss_fwd(_lib.ptr(a), _lib.ptr(b), n, _LOSS[loss], _TONEMAP[tonemapper],
                                               _lib.ptr(partial), _lib.current_stream(a.device)), "gsb_image_loss_fwd")
        ctx.save_for_backward(a, b)
        ctx.meta = (loss, tonemapper, npix, img.shape, target.shape)
        return partial.sum() / npix          # reference ops.py:497

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        loss, tonemapper, npix, s_img, s_tgt = ctx.meta
        g = g.float().contiguous().reshape(1)
        gi = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gt = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.lib.gsb_image_loss_bwd(_lib.ptr(a), _lib.ptr(b), a.numel(), _LOSS[loss], _TONEMAP[tonemapper],
                                               _lib.ptr(g), 1.0 / npix, _lib.ptr(gi), _lib.ptr(gt),
                                               _lib.current_stream(a.device)), "gsb_image_loss_bwd")
        return (None if gi is None else gi.sum_to_size(s_img), None if gt is None else gt.sum_to_size(s_tgt), None, None)


def image_loss(img, target, loss="l1", tonemapper="none", use_python=False):
    """Scalar HDR image loss (reference ops.py:479-503); loss in {l1,mse,smape,relmse}, tonemapper in {none,log_srgb}."""
    _no_python(use_python)
    _need_cuda(img, "image_loss")
    if loss not in _LOSS or tonemapper not in _TONEMAP:
        raise ValueError(f"image_loss: unknown loss/tonemapper {loss}/{tonemapper}")
    return _ImageLoss.apply(img, target, loss, tonemapper)


# ------------------------------------------------------------------------------------------------
# Pointwise BSDF operators (csrc/bsdf_ops.cu): the rest of the names the reference's package exports
# (renderutils/__init__.py:10-11).  One generic autograd node: the inputs are broadcast over their leading
# dimensions OUTSIDE the node (differentiable expand, so autograd reduces the gradients of broadcast inputs),
# the node itself sees dense [n, C] arrays.
class _Pointwise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, name, out_channels, extra, *ins):
        n = ins[0].numel() // ins[0].shape[-1]
        out = torch.empty((*ins[0].shape[:-1], out_channels), dtype=torch.float32, device=ins[0].device)
        fn = getattr(_lib.lib, f"gsb_{name}_fwd")
        _lib.check(fn(*[_lib.ptr(t) for t in ins], *extra, n, _lib.ptr(out), _lib.current_stream(out.device)), f"gsb_{name}_fwd")
        ctx.save_for_backward(*ins)
        ctx.meta = (name, extra, n)
        return out

    @staticmethod
    def backward(ctx, g_out):
        ins = ctx.saved_tensors
        name, extra, n = ctx.meta
        g = g_out.float().contiguous()
        grads = [torch.empty_like(t) for t in ins]
        fn = getattr(_lib.lib, f"gsb_{name}_bwd")
        _lib.check(fn(*[_lib.ptr(t) for t in ins], *extra, _lib.ptr(g), n, *[_lib.ptr(t) for t in grads], _lib.current_stream(g.device)),
                   f"gsb_{name}_bwd")
        return (None, None, None, *grads)


class _PbrBsdf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, min_roughness, bsdf, *ins):
        n = ins[0].numel() // 3
        out = torch.empty_like(ins[0])
        ptrs = (ctypes.c_void_p * 6)(*[_lib.ptr(t) for t in ins])
        _lib.check(_lib.lib.gsb_pbr_bsdf_fwd(ptrs, min_roughness, bsdf, n, _lib.ptr(out), _lib.current_stream(out.device)), "gsb_pbr_bsdf_fwd")
        ctx.save_for_backward(*ins)
        ctx.meta = (min_roughness, bsdf, n)
        return out

    @staticmethod
    def backward(ctx, g_out):
        ins = ctx.saved_tensors
        min_roughness, bsdf, n = ctx.meta
        g = g_out.float().contiguous()
        grads = [torch.empty_like(t) for t in ins]
        ptrs = (ctypes.c_void_p * 6)(*[_lib.ptr(t) for t in ins])
        gptrs = (ctypes.c_void_p * 6)(*[_lib.ptr(t) for t in grads])
        _lib.check(_lib.lib.gsb_pbr_bsdf_bwd(ptrs, min_roughness, bsdf, _lib.ptr(g), n, gptrs, _lib.current_stream(g.device)), "gsb_pbr_bsdf_bwd")
        return (None, None, *grads)


def _dense(what, channels, *tensors):
    """The operands broadcast over their leading dimensions, fp32, contiguous; scalars (the reference passes e.g. f90 = 1) become
    one-element tensors first.  `channels[k]` is operand k's trailing size (1 or 3)."""
    ref = next(t for t in tensors if torch.is_tensor(t))
    _need_cuda(ref, what)
    ts = [t if torch.is_tensor(t) else torch.full((1,), float(t), dtype=torch.float32, device=ref.device) for t in tensors]
    lead = torch.broadcast_shapes(*[t.shape[:-1] for t in ts])
    out = []
    for t, c in zip(ts, channels):
        if t.shape[-1] not in (1, c):
            raise RuntimeError(f"{what}: trailing dimension {t.shape[-1]} where {c} is expected")
        out.append(t.float().expand(*lead, c).contiguous())
    return out


def _fresnel_shlick(f0, f90, cosTheta, use_python=False):
    """f0 + (f90 - f0) (1 - clamp(cos))^5 (reference ops.py:91-113)."""
    _no_python(use_python)
    return _Pointwise.apply("fresnel_shlick", 3, (), *_dense("_fresnel_shlick", (3, 3, 1), f0, f90, cosTheta))


def _ndf_ggx(alphaSqr, cosTheta, use_python=False):
    """GGX normal distribution (reference ops.py:116-137)."""
    _no_python(use_python)
    return _Pointwise.apply("ndf_ggx", 1, (), *_dense("_ndf_ggx", (1, 1), alphaSqr, cosTheta))


def _lambda_ggx(alphaSqr, cosTheta, use_python=False):
    """Smith Lambda of GGX (reference ops.py:139-160)."""
    _no_python(use_python)
    return _Pointwise.apply("lambda_ggx", 1, (), *_dense("_lambda_ggx", (1, 1), alphaSqr, cosTheta))


def _masking_smith(alphaSqr, cosThetaI, cosThetaO, use_python=False):
    """Height-correlated Smith masking-shadowing (reference ops.py:162-183)."""
    _no_python(use_python)
    return _Pointwise.apply("masking_smith", 1, (), *_dense("_masking_smith", (1, 1, 1), alphaSqr, cosThetaI, cosThetaO))


def lambert(nrm, wi, use_python=False):
    """Lambertian BSDF max(n.wi, 0) / pi -> [..., 1] (reference ops.py:251-271)."""
    _no_python(use_python)
    return _Pointwise.apply("lambert", 1, (), *_dense("lambert", (3, 3), nrm, wi))


def frostbite_diffuse(nrm, wi, wo, linearRoughness, use_python=False):
    """Frostbite's normalised Disney diffuse -> [..., 1] (reference ops.py:285-307)."""
    _no_python(use_python)
    return _Pointwise.apply("frostbite", 1, (), *_dense("frostbite_diffuse", (3, 3, 3, 1), nrm, wi, wo, linearRoughness))


def pbr_specular(col, nrm, wo, wi, alpha, min_roughness=0.08, use_python=False):
    """GGX specular lobe -> [..., 3]; alpha [..., 1] (reference ops.py:322-347)."""
    _no_python(use_python)
    return _Pointwise.apply("pbr_specular", 3, (float(min_roughness),), *_dense("pbr_specular", (3, 3, 3, 3, 1), col, nrm, wo, wi, alpha))


def pbr_bsdf(kd, arm, pos, nrm, view_pos, light_pos, min_roughness=0.08, bsdf="lambert", use_python=False):
    """Diffuse ('lambert' or 'frostbite') + GGX specular of a metallic-roughness material -> [..., 3] (reference ops.py:363-393)."""
    _no_python(use_python)
    return _PbrBsdf.apply(float(min_roughness), 1 if bsdf == "frostbite" else 0,
                          *_dense("pbr_bsdf", (3,) * 6, kd, arm, pos, nrm, view_pos, light_pos))


class _XfmVectors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vectors, matrix):
        v = vectors.detach().float().contiguous()
        m = matrix.detach().float().contiguous()
        B, N = m.shape[0], v.shape[1]
        batched = 1 if v.shape[0] != 1 else 0
        if batched and v.shape[0] != B:
            raise RuntimeError("xfm_vectors: vectors batch must be 1 or match the matrix batch")
        out = torch.empty((B, N, 3), dtype=torch.float32, device=v.device)
        _lib.check(_lib.lib.gsb_xfm_vectors_fwd(_lib.ptr(v), _lib.ptr(m), B, N, batched, _lib.ptr(out), _lib.current_stream(v.device)),
                   "gsb_xfm_vectors_fwd")
        ctx.save_for_backward(m)
        ctx.meta = (B, N, batched, vectors.shape)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (m,) = ctx.saved_tensors
        B, N, batched, shape = ctx.meta
        g = g_out.float().contiguous()
        g_vec = torch.empty(shape, dtype=torch.float32, device=g.device)
        _lib.check(_lib.lib.gsb_xfm_vectors_bwd(_lib.ptr(m), _lib.ptr(g), B, N, batched, _lib.ptr(g_vec), _lib.current_stream(g.device)),
                   "gsb_xfm_vectors_bwd")
        return g_vec, None


def xfm_vectors(vectors, matrix, use_python=False):
    """[1|B,N,3] x [B,4,4] -> [B,N,3], the rotation / scale part only (w = 0; reference ops.py:540-556)."""
    _no_python(use_python)
    _need_cuda(vectors, "xfm_vectors")
    return _XfmVectors.apply(vectors, matrix)
