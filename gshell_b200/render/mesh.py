"""Mesh container + smooth vertex normals (reference render/mesh.py:20-74, :212-237)."""
import torch

from .. import _lib


class Mesh:
    def __init__(self, v_pos=None, t_pos_idx=None, v_nrm=None, t_nrm_idx=None, v_tex=None, t_tex_idx=None, v_tng=None,
                 t_tng_idx=None, material=None, base=None):
        self.v_pos, self.v_nrm, self.v_tex, self.v_tng = v_pos, v_nrm, v_tex, v_tng
        self.t_pos_idx, self.t_nrm_idx, self.t_tex_idx, self.t_tng_idx = t_pos_idx, t_nrm_idx, t_tex_idx, t_tng_idx
        self.material = material
        if base is not None:
            self.copy_none(base)

    def copy_none(self, other):
        for k in ("v_pos", "t_pos_idx", "v_nrm", "t_nrm_idx", "v_tex", "t_tex_idx", "v_tng", "t_tng_idx", "material"):
            if getattr(self, k) is None:
                setattr(self, k, getattr(other, k))

    def clone(self):
        out = Mesh(base=self)
        for k in ("v_pos", "t_pos_idx", "v_nrm", "t_nrm_idx", "v_tex", "t_tex_idx", "v_tng", "t_tng_idx"):
            v = getattr(out, k)
            if v is not None:
                setattr(out, k, v.clone().detach())
        return out


class _VertexNormals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_pos, tris):
        v = v_pos.detach().float().contiguous()
        t = tris.int().contiguous()
        acc = torch.empty_like(v)
        nrm = torch.empty_like(v)
        _lib.check(_lib.lib.gsb_vertex_normals_fwd(_lib.ptr(v), _lib.ptr(t), v.shape[0], t.shape[0], _lib.ptr(acc),
                                                   _lib.ptr(nrm), _lib.current_stream(v.device)), "gsb_vertex_normals_fwd")
        ctx.save_for_backward(v, t, acc)
        return nrm

    @staticmethod
    def backward(ctx, g):
        v, t, acc = ctx.saved_tensors
        g = g.float().contiguous()
        g_acc = torch.empty_like(v)
        g_v = torch.empty_like(v)
        _lib.check(_lib.lib.gsb_vertex_normals_bwd(_lib.ptr(v), _lib.ptr(t), _lib.ptr(acc), _lib.ptr(g), v.shape[0],
                                                   t.shape[0], _lib.ptr(g_acc), _lib.ptr(g_v),
                                                   _lib.current_stream(v.device)), "gsb_vertex_normals_bwd")
        return g_v, None


def vertex_normals(v_pos, t_pos_idx):
    return _VertexNormals.apply(v_pos, t_pos_idx)


def auto_normals(imesh):
    """Reference mesh.py:212-237."""
    if imesh.v_pos.shape[0] == 0:
        return Mesh(v_nrm=torch.zeros_like(imesh.v_pos), t_nrm_idx=imesh.t_pos_idx, base=imesh)
    return Mesh(v_nrm=vertex_normals(imesh.v_pos, imesh.t_pos_idx), t_nrm_idx=imesh.t_pos_idx, base=imesh)


def compute_tangents(imesh, v_tng=None):
    """Reference mesh.py:243-247: adopt a given per-vertex tangent field (normalised, made perpendicular to the vertex
    normals).  The uv-based construction (:249-286) is only reached by meshes with texture coordinates, which the G-Shell path
    never builds."""
    if v_tng is None:
        raise NotImplementedError("tangents from texture coordinates: not on the G-Shell path (no uv atlas is built)")
    from . import util
    v_tng = util.safe_normalize(v_tng)
    v_tng = util.safe_normalize(v_tng - util.dot(v_tng, imesh.v_nrm) * imesh.v_nrm)
    return Mesh(v_tng=v_tng, t_tng_idx=imesh.t_nrm_idx, base=imesh)

