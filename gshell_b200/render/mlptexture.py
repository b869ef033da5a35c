"""`MLPTexture3D` -- the reference's learned material field (render/mlptexture.py:49-106): a multiresolution hash-grid
encoding of the surface position followed by a small bias-free ReLU MLP and a sigmoid range map; `material['kd_ks'].sample(gb_pos)`
in `render.shade` (reference render.py:69).  Same constructor and methods (`sample`, `clamp_`, `cleanup`).

The MLP is plain PyTorch in the reference too (`_MLP`, :18-43; cuBLAS GEMMs of width 32).  The encoding is tiny-cuda-nn's
`HashGrid` there -- third-party and absent here -- and `csrc/hashgrid.cu` here: an own implementation of the published algorithm
in the reference's configuration (16 levels x 2 features, 2^19 entries, 16 -> 4096), fp32 table; parity with tcnn is unpinned
(upstream of the measured hot path)."""
import ctypes
import math

import numpy as np
import torch

from .. import _lib


def hashgrid_levels(n_levels=16, base_resolution=16, per_level_scale=None, log2_hashmap_size=19, desired_resolution=4096):
    """Level layout of the table: (offsets uint32[L+1] in entries, resolutions uint32[L], scales float32[L]):
    scale_l = base * s^l - 1 (evaluated in double, snapped to the integer it misses by rounding, stored as float32),
    resolution_l = ceil(scale_l) + 1, entries_l = min(round_up(resolution_l^3, 8), 2^log2_hashmap_size)."""
    if per_level_scale is None:
        per_level_scale = math.exp(math.log(desired_resolution / base_resolution) / (n_levels - 1))
    offs = np.zeros(n_levels + 1, dtype=np.uint32)
    ress = np.zeros(n_levels, dtype=np.uint32)
    scales = np.zeros(n_levels, dtype=np.float32)
    for lvl in range(n_levels):
        scale = base_resolution * per_level_scale ** lvl - 1.0
        if abs(scale - round(scale)) < 1e-9 * max(1.0, abs(scale)):
            scale = float(round(scale))
        scale = float(np.float32(scale))
        res = int(math.ceil(scale)) + 1
        n = min((res ** 3 + 7) // 8 * 8, 1 << log2_hashmap_size)
        offs[lvl + 1] = offs[lvl] + n
        ress[lvl] = res
        scales[lvl] = scale
    return offs, ress, scales


class _HashGrid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x01, table, layout):
        offs, ress, scales = layout
        _lib.require_cuda(x01, "gshell_b200.render.mlptexture")
        x = x01.detach().float().contiguous()
        t = table.detach().float().contiguous()
        n, L = x.shape[0], len(ress)
        out = torch.empty((n, 2 * L), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib.gsb_hashgrid_fwd(_lib.ptr(x), n, _lib.ptr(t), offs.ctypes.data_as(ctypes.c_void_p), ress.ctypes.data_as(ctypes.c_void_p),
                                             scales.ctypes.data_as(ctypes.c_void_p), L, _lib.ptr(out), _lib.current_stream(x.device)),
                   "gsb_hashgrid_fwd")
        ctx.save_for_backward(x, t)
        ctx.layout = layout
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, t = ctx.saved_tensors
        offs, ress, scales = ctx.layout
        g = g_out.float().contiguous()
        g_table = torch.empty_like(t) if ctx.needs_input_grad[1] else None
        g_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        _lib.check(_lib.lib.gsb_hashgrid_bwd(_lib.ptr(x), x.shape[0], _lib.ptr(t), offs.ctypes.data_as(ctypes.c_void_p),
                                             ress.ctypes.data_as(ctypes.c_void_p), scales.ctypes.data_as(ctypes.c_void_p), len(ress),
                                             _lib.ptr(g), _lib.ptr(g_table), _lib.ptr(g_x), _lib.current_stream(x.device)),
                   "gsb_hashgrid_bwd")
        return g_x, g_table, None


class _ScaleGrad(torch.autograd.Function):
    """identity whose gradient is multiplied by `k` (the reference's two backward hooks, mlptexture.py:30,76)"""

    @staticmethod
    def forward(ctx, x, k):
        ctx.k = k
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.k, None


class HashGridEncoding(torch.nn.Module):
    """`tcnn.Encoding(3, {"otype": "HashGrid", ...})` as the reference configures it: input [N,3] in [0,1], output [N, 2 * n_levels]."""

    def __init__(self, n_levels=16, base_resolution=16, per_level_scale=None, log2_hashmap_size=19, desired_resolution=4096, device="cuda"):
        super().__init__()
        self.layout = hashgrid_levels(n_levels, base_resolution, per_level_scale, log2_hashmap_size, desired_resolution)
        self.n_output_dims = 2 * n_levels
        # the usual initialisation of this encoding: U(-1e-4, 1e-4)
        self.params = torch.nn.Parameter((torch.rand((int(self.layout[0][-1]), 2), device=device) * 2 - 1) * 1e-4)

    def forward(self, x01):
        return _HashGrid.apply(x01, self.params, self.layout)


class _MLP(torch.nn.Module):
    """bias-free Linear / ReLU stack, Kaiming-uniform weights (reference :18-43)"""

    def __init__(self, cfg, device="cuda"):
        super().__init__()
        dims = [cfg["n_input_dims"]] + [cfg["n_neurons"]] * cfg["n_hidden_layers"] + [cfg["n_output_dims"]]
        layers = []
        for i in range(len(dims) - 1):
            lin = torch.nn.Linear(dims[i], dims[i + 1], bias=False)
            torch.nn.init.kaiming_uniform_(lin.weight, nonlinearity="relu")
            layers.append(lin)
            if i < len(dims) - 2:
                layers.append(torch.nn.ReLU())
        self.net = torch.nn.Sequential(*layers).to(device)

    def forward(self, x):
        return self.net(x.to(torch.float32))


class MLPTexture3D(torch.nn.Module):
    def __init__(self, AABB, channels=3, internal_dims=32, hidden=2, min_max=None, use_float16=False):
        super().__init__()
        self.channels = channels
        self.internal_dims = internal_dims
        self.AABB = AABB
        self.min_max = min_max
        self.use_float16 = use_float16
        dev = AABB.device if torch.is_tensor(AABB) else "cuda"
        self.encoder = HashGridEncoding(device=dev)
        self.gradient_scaling = 128.0
        self.net = _MLP({"n_input_dims": self.encoder.n_output_dims, "n_output_dims": channels, "n_hidden_layers": hidden,
                         "n_neurons": internal_dims}, device=dev)

    def _fused_inference_ok(self):
        """the one-launch inference kernel covers the reference's own configuration: 16 levels x 2 features, two hidden layers of 32"""
        lin = [m for m in self.net.net if isinstance(m, torch.nn.Linear)]
        return (not self.use_float16 and self.encoder.n_output_dims == 32 and len(lin) == 3 and self.channels <= 16
                and all(m.bias is None for m in lin) and [tuple(m.weight.shape) for m in lin] == [(32, 32), (32, 32), (self.channels, 32)])

    @torch.no_grad()
    def _sample_fused(self, texc):
        """Encoding, MLP and range map in ONE kernel (csrc/hashgrid.cu::k_field_infer): no [n, 32] intermediate reaches HBM."""
        _lib.require_cuda(texc, "gshell_b200.render.mlptexture")
        pos = texc.detach().float().reshape(-1, 3).contiguous()
        offs, ress, scales = self.encoder.layout
        w1, w2, w3 = (m.weight.detach().float().contiguous() for m in self.net.net if isinstance(m, torch.nn.Linear))
        table = self.encoder.params.detach().float().contiguous()
        aabb = np.ascontiguousarray(torch.cat([self.AABB[0].reshape(3), self.AABB[1].reshape(3)]).detach().cpu().numpy(), dtype=np.float32)
        rng = np.ascontiguousarray(torch.cat([self.min_max[0].reshape(-1), self.min_max[1].reshape(-1)]).detach().cpu().numpy(), dtype=np.float32)
        out = torch.empty((pos.shape[0], self.channels), dtype=torch.float32, device=pos.device)
        as_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)          # noqa: E731
        _lib.check(_lib.lib.gsb_field_infer(_lib.ptr(pos), pos.shape[0], _lib.ptr(table), as_p(offs), as_p(ress), as_p(scales), len(ress),
                                            _lib.ptr(w1), _lib.ptr(w2), _lib.ptr(w3), self.channels, as_p(aabb), as_p(rng), _lib.ptr(out),
                                            _lib.current_stream(pos.device)), "gsb_field_infer")
        return out.view(*texc.shape[:-1], self.channels)

    def sample(self, texc):
        """texc [..., 3] world positions -> [..., channels] within min_max (reference :86-98).  Without autograd (validation renders,
        texture baking) the whole field is one kernel; with autograd the encoding kernel feeds the PyTorch MLP as in the reference."""
        if not torch.is_grad_enabled() and self._fused_inference_ok():
            return self._sample_fused(texc)
        lo, hi = self.AABB[0], self.AABB[1]
        x = torch.clamp((texc.reshape(-1, 3) - lo[None]) / (hi[None] - lo[None]), min=0, max=1)
        # reference hooks: the encoder's parameters see the gradient x128, its input sees it /128 x128 = unchanged
        p_enc = self.encoder(_ScaleGrad.apply(x.contiguous(), 1.0 / self.gradient_scaling))
        p_enc = _ScaleGrad.apply(p_enc, self.gradient_scaling)
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.use_float16):
            out = self.net(p_enc)
        out = torch.sigmoid(out.float()) * (self.min_max[1][None, :] - self.min_max[0][None, :]) + self.min_max[0][None, :]
        return out.view(*texc.shape[:-1], self.channels)

    def clamp_(self):
        pass

    def cleanup(self):
        pass
