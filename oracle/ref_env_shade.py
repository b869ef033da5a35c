"""TEST INFRASTRUCTURE: Python face of oracle/_ref/libref_env_shade.so -- the reference's own integrator
(render/optixutils/c_src/envsampling/kernel.cu, unmodified) compiled for the CPU by oracle/build_ref.py.  Same arguments as the
reference's `env_shade_fwd / env_shade_bwd` (torch_bindings.cpp:123-272), plus the occluder mesh that its OptiX GAS would hold.
Used to pin oracle/shade_oracle.py::env_shade; never imported by the product."""
import contextlib
import ctypes

import numpy as np
import torch

from . import build_ref

_libs = {}
_variant = ""


def lib():
    if _variant not in _libs:
        path = build_ref.build(variant=_variant)
        if path is None:
            raise RuntimeError(f"oracle/_ref/libref_env_shade{'_' + _variant if _variant else ''}.so is missing and the reference "
                               "checkout is not mounted")
        handle = ctypes.CDLL(path)
        handle.ref_env_shade.restype = None
        _libs[_variant] = handle
    return _libs[_variant]


@contextlib.contextmanager
def code_generation(variant):
    """Run the calls inside through another compilation of the same unmodified source (build_ref.VARIANTS: "fma", "fast")."""
    global _variant
    prev, _variant = _variant, variant
    try:
        yield
    finally:
        _variant = prev


def _f(t, shape=None):
    t = t.detach().float().cpu()
    if shape is not None:
        t = t.expand(shape)
    return np.ascontiguousarray(t.numpy())


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _common(mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols, perms, verts, tris):
    B, H, W = mask.shape
    full = (B, H, W, 3)
    a = dict(mask=_f(mask), ro=_f(ro, full), pos=_f(gb_pos, full), nrm=_f(gb_normal, full),
             vpos=_f(gb_view_pos.reshape(-1, 1, 1, 3), (B, 1, 1, 3)), kd=_f(gb_kd, full), ks=_f(gb_ks, full), light=_f(light),
             pdf=_f(pdf), rows=_f(rows.reshape(-1)), cols=_f(cols),
             perms=np.ascontiguousarray(perms.detach().cpu().numpy().astype(np.int32)))
    a["verts"] = None if verts is None else _f(verts.reshape(-1, 3))
    a["tris"] = None if tris is None else np.ascontiguousarray(tris.detach().cpu().numpy().astype(np.int32).reshape(-1, 3))
    return a, (B, H, W)


def _call(backward, a, dims, bsdf, n, seed, shadow_scale, outs):
    B, H, W = dims
    lh, lw = a["light"].shape[0], a["light"].shape[1]
    n_tris = 0 if a["tris"] is None else a["tris"].shape[0]
    lib().ref_env_shade(ctypes.c_int(backward), B, H, W, _p(a["mask"]), _p(a["ro"]), _p(a["pos"]), _p(a["nrm"]), _p(a["vpos"]),
                        _p(a["kd"]), _p(a["ks"]), _p(a["light"]), _p(a["pdf"]), _p(a["rows"]), _p(a["cols"]), _p(a["perms"]),
                        a["perms"].shape[0], lh, lw, ctypes.c_uint(bsdf), ctypes.c_uint(n), ctypes.c_uint(seed & 0xFFFFFFFF),
                        ctypes.c_float(shadow_scale), _p(a["verts"]), _p(a["tris"]), n_tris, *[_p(o) for o in outs])


def env_shade_fwd(mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols, perms, bsdf=0, n_samples_x=2,
                  rnd_seed=0, shadow_scale=0.0, verts=None, tris=None):
    a, dims = _common(mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols, perms, verts, tris)
    diff, spec = (np.empty(dims + (3,), np.float32) for _ in range(2))
    _call(0, a, dims, bsdf, n_samples_x, rnd_seed, shadow_scale, [diff, spec, None, None, None, None, None, None, None])
    return torch.from_numpy(diff), torch.from_numpy(spec)


def env_shade_bwd(mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols, perms, diff_grad, spec_grad,
                  bsdf=0, n_samples_x=2, rnd_seed=0, shadow_scale=0.0, verts=None, tris=None):
    """Returns (d_pos, d_normal, d_kd, d_ks, d_light) like env_shade_bwd of the reference."""
    a, dims = _common(mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols, perms, verts, tris)
    full = dims + (3,)
    gd, gs = _f(diff_grad, full), _f(spec_grad, full)
    outs = [np.empty(full, np.float32) for _ in range(4)] + [np.empty(a["light"].shape, np.float32)]
    _call(1, a, dims, bsdf, n_samples_x, rnd_seed, shadow_scale, [None, None, gd, gs] + outs)
    return tuple(torch.from_numpy(o) for o in outs)
