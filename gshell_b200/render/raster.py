"""Rasterise / interpolate autograd operators over csrc/raster.cu.  They stand where the reference calls
nvdiffrast (`dr.DepthPeeler(...).rasterize_next_layer()` render.py:377-379, `dr.interpolate` render.py:26),
with nvdiffrast's tensor conventions so `render.py` reads the same.  `antialias` (csrc/antialias.cu) follows
nvdiffrast's documented algorithm; like the rasteriser it is an own implementation, parity unpinned."""
import torch

from .. import _lib


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, clip, tris, resolution):
        c = clip.detach().float().contiguous()
        t = tris.int().contiguous()
        H, W = int(resolution[0]), int(resolution[1])
        B, V = c.shape[0], c.shape[1]
        dev = c.device
        rast = torch.empty((B, H, W, 4), dtype=torch.float32, device=dev)
        db = torch.empty((B, H, W, 4), dtype=torch.float32, device=dev)
        zbuf = torch.empty((B * H * W,), dtype=torch.int64, device=dev)
        _lib.check(_lib.lib.gsb_rasterize_fwd(_lib.ptr(c), _lib.ptr(t), B, V, t.shape[0], 1, H, W, _lib.ptr(zbuf),
                                              _lib.ptr(rast), _lib.ptr(db), _lib.current_stream(dev)), "gsb_rasterize_fwd")
        ctx.save_for_backward(c, t, rast)
        ctx.mark_non_differentiable(db)
        return rast, db

    @staticmethod
    def backward(ctx, g_rast, _g_db):
        c, t, rast = ctx.saved_tensors
        B, V = c.shape[0], c.shape[1]
        H, W = rast.shape[1], rast.shape[2]
        g = g_rast.float().contiguous()
        g_clip = torch.zeros_like(c)
        _lib.check(_lib.lib.gsb_rasterize_bwd(_lib.ptr(c), _lib.ptr(t), _lib.ptr(rast), _lib.ptr(g), B, V, 1, H, W,
                                              _lib.ptr(g_clip), _lib.current_stream(c.device)), "gsb_rasterize_bwd")
        return g_clip, None, None


def rasterize(clip, tris, resolution):
    """clip [B,V,4], tris int [F,3], resolution (H,W) -> (rast [B,H,W,4] = (u,v,z/w,id+1), rast_db [B,H,W,4])."""
    _lib.require_cuda(clip, "rasterize")
    return _Rasterize.apply(clip, tris, resolution)


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tris, rast_db):
        a = attr.detach().float().contiguous()
        r = rast.detach().float().contiguous()
        t = tris.int().contiguous()
        B, H, W, _ = r.shape
        V, C = a.shape[1], a.shape[2]
        batched = 1 if a.shape[0] != 1 else 0
        dev = a.device
        out = torch.empty((B, H, W, C), dtype=torch.float32, device=dev)
        out_d = None
        db = None
        if rast_db is not None:
            db = rast_db.detach().float().contiguous()
            out_d = torch.empty((B, H, W, C * 2), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib.gsb_interpolate_fwd(_lib.ptr(a), _lib.ptr(t), _lib.ptr(r), _lib.ptr(db), B, V, C, batched, H, W,
                                                _lib.ptr(out), _lib.ptr(out_d), _lib.current_stream(dev)),
                   "gsb_interpolate_fwd")
        ctx.save_for_backward(a, r, t)
        ctx.batched = batched
        if out_d is None:
            out_d = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(out_d)
        return out, out_d

    @staticmethod
    def backward(ctx, g_out, _g_d):
        a, r, t = ctx.saved_tensors
        B, H, W, _ = r.shape
        V, C = a.shape[1], a.shape[2]
        g = g_out.float().contiguous()
        need_attr, need_rast = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_attr = torch.zeros_like(a) if need_attr else None
        g_rast = torch.empty_like(r) if need_rast else None
        _lib.check(_lib.lib.gsb_interpolate_bwd(_lib.ptr(a), _lib.ptr(t), _lib.ptr(r), _lib.ptr(g), B, V, C, ctx.batched, H, W,
                                                _lib.ptr(g_attr), _lib.ptr(g_rast), _lib.current_stream(a.device)),
                   "gsb_interpolate_bwd")
        return g_attr, g_rast, None, None


def interpolate(attr, rast, tris, rast_db=None):
    """attr [1|B,V,C] -> ([B,H,W,C], derivatives [B,H,W,2C] or empty).  Same contract as dr.interpolate with
    diff_attrs='all' when rast_db is given (the derivative layout is (dX,dY) interleaved per channel)."""
    return _Interpolate.apply(attr, rast, tris, rast_db)


class _Analysis:
    """Silhouette work items of one frame (csrc/antialias.cu::gsb_antialias_analyse), shared by every buffer antialiased
    against the same (rast, clip, tris)."""

    def __init__(self, rast, clip, tris):
        self.rast, self.clip, self.tris = rast, clip, tris           # kept alive: the cache key below is their addresses
        B, H, W, _ = rast.shape
        dev = rast.device
        L = _lib.lib
        self.cap = 2 * B * H * W
        self.items = torch.empty(self.cap * int(L.gsb_antialias_item_bytes()), dtype=torch.uint8, device=dev)
        self.n_items = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(int(L.gsb_antialias_hash_slots(tris.shape[0])) * 16, dtype=torch.uint8, device=dev)
        _lib.check(L.gsb_antialias_analyse(_lib.ptr(rast), _lib.ptr(clip), _lib.ptr(tris), B, H, W, clip.shape[1], tris.shape[0], _lib.ptr(ws),
                                           _lib.ptr(self.items), _lib.ptr(self.n_items), self.cap, _lib.current_stream(dev)),
                   "gsb_antialias_analyse")


_last_analysis = None


def _analysis(rast, clip, tris):
    global _last_analysis
    key = (rast.data_ptr(), rast._version, clip.data_ptr(), clip._version, tris.data_ptr(), tuple(rast.shape))
    if _last_analysis is None or _last_analysis[0] != key:
        _last_analysis = (key, _Analysis(rast, clip, tris))
    return _last_analysis[1]


class _Antialias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, pos_clip, an):
        c = color.detach().float().contiguous()
        out = c.clone()
        _lib.check(_lib.lib.gsb_antialias_fwd(_lib.ptr(c), _lib.ptr(an.items), _lib.ptr(an.n_items), an.cap, c.shape[-1], _lib.ptr(out),
                                              _lib.current_stream(c.device)), "gsb_antialias_fwd")
        ctx.save_for_backward(c)
        ctx.an = an
        return out

    @staticmethod
    def backward(ctx, g_out):
        (c,) = ctx.saved_tensors
        an = ctx.an
        g = g_out.float().contiguous()
        g_color = g.clone() if ctx.needs_input_grad[0] else None
        g_clip = torch.zeros_like(an.clip) if ctx.needs_input_grad[1] else None
        B, H, W, _ = an.rast.shape
        _lib.check(_lib.lib.gsb_antialias_bwd(_lib.ptr(c), _lib.ptr(g), _lib.ptr(an.items), _lib.ptr(an.n_items), an.cap, c.shape[-1],
                                              _lib.ptr(an.clip), an.clip.shape[1], H, W, _lib.ptr(g_color), _lib.ptr(g_clip),
                                              _lib.current_stream(c.device)), "gsb_antialias_bwd")
        return g_color, g_clip, None


def antialias(color, rast, pos_clip, tris):
    """dr.antialias(color, rast, pos_clip, tris) (reference render.py:358): blends colours across silhouette edges by the edge's
    sub-pixel position and carries the gradient of every blended channel -- coverage / alpha included -- to the clip-space
    vertex positions.  Own implementation (csrc/antialias.cu); nvdiffrast itself is not available: parity unpinned."""
    _lib.require_cuda(color, "antialias")
    an = _analysis(rast.detach().float().contiguous(), pos_clip.detach().float().contiguous(), tris.int().contiguous())
    return _Antialias.apply(color, pos_clip, an)
