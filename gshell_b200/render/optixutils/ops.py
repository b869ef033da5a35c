"""Environment-light MC shading and the bilateral denoiser on plain CUDA (csrc/env_shade.cu,
csrc/denoise.cu), behind the names of the reference's render/optixutils/ops.py.

Differences a maintainer should know (all documented in DESIGN.md):
  * no OptiX / NVRTC: the module imports without a driver-side libnvoptix and never JIT-builds;
  * `env_shade` does not synchronise the stream nor allocate per call (reference torch_bindings.cpp:174-185);
  * shadow rays are traced against a uniform-grid occluder rebuilt by `optix_build_bvh` each iteration
    (csrc/occluder.cu: count -> scan -> fill; bricks -> cells -> sub-voxel bits, csrc/trace_core.cuh) instead of an
    OptiX GAS; a context without a mesh means "nothing occludes".
"""
import numpy as np
import torch

from ... import _lib

# grid cells per mesh face (R = cbrt(this * F)); 2 measured best on the probes of profiles/prof_shadow.py
import os as _os
OCCLUDER_CELLS_PER_FACE = float(_os.environ.get("GSB_OCC_CELLS_PER_FACE", "2.0"))

_BSDF = ["pbr", "diffuse", "white"]      # order = the kernel's BSDF ids (reference ops.py:142)


class OptiXContext:
    """Occluder state of the current iteration (the reference wraps an OptiX pipeline + GAS here, ops.py:128-131)."""

    def __init__(self):
        self.verts = None
        self.tris = None
        self.occluder = None      # device struct (uint8 tensor) handed to the integrator, or None
        self._keep = None         # tensors the struct points into

    def bvh_ptr(self):
        return None if self.occluder is None else self.occluder.data_ptr()


def optix_build_bvh(optix_ctx, verts, tris, rebuild):
    """Reference ops.py:133-139 (`rebuild` is accepted for signature parity; the grid is always rebuilt -- the
    reference's only caller passes rebuild=1, gshell_tets_geometry.py:211)."""
    L = _lib.lib
    v = verts.detach().reshape(-1, 3).float().contiguous()
    t = tris.reshape(-1, 3).int().contiguous()
    optix_ctx.verts, optix_ctx.tris = v, t
    F = t.shape[0]
    if F == 0 or v.shape[0] == 0:
        optix_ctx.occluder, optix_ctx._keep = None, None
        return
    dev = v.device
    stream = _lib.current_stream(dev)
    R = int(min(320, max(4, round(OCCLUDER_CELLS_PER_FACE ** (1.0 / 3.0) * F ** (1.0 / 3.0)))))
    n_cells = int(L.gsb_occluder_cells(R))                           # padded to whole 4x4x4 bricks, brick-major order
    lo, hi = torch.aminmax(v, dim=0)
    lo, hi = lo.contiguous(), hi.contiguous()
    occ = torch.empty(int(L.gsb_occluder_struct_bytes()), dtype=torch.uint8, device=dev)
    cell_start = torch.empty(n_cells + 1, dtype=torch.int32, device=dev)
    scan_ws = torch.empty(int(L.gsb_occluder_scan_ws_ints(n_cells)), dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    brick_bits = torch.empty(int(L.gsb_occluder_brick_words(R)), dtype=torch.int64, device=dev)
    _lib.check(L.gsb_occluder_build_count(_lib.ptr(v), _lib.ptr(t), F, _lib.ptr(lo), _lib.ptr(hi), R, _lib.ptr(occ),
                                          _lib.ptr(cell_start), _lib.ptr(scan_ws), _lib.ptr(brick_bits), _lib.ptr(total), stream),
               "gsb_occluder_build_count")
    n_entries = int(total.item())                                    # one host read: sizes the entry list
    cell_tris = torch.empty((n_entries + 1, 12), dtype=torch.float32, device=dev)     # (v0,e1,e2) per (cell, triangle)
    cursor = torch.empty(n_cells, dtype=torch.int32, device=dev)
    cell_recs = torch.empty((n_cells, 4), dtype=torch.int32, device=dev)                # {first entry, entries, sub-voxel bits}
    _lib.check(L.gsb_occluder_build_fill(_lib.ptr(v), _lib.ptr(t), F, R, _lib.ptr(occ), _lib.ptr(cell_start), _lib.ptr(cursor),
                                         _lib.ptr(cell_recs), _lib.ptr(cell_tris), stream), "gsb_occluder_build_fill")
    optix_ctx.occluder = occ
    optix_ctx._keep = (cell_recs, cell_tris, brick_bits)
    optix_ctx.grid_res, optix_ctx.n_entries = R, n_entries


# HBM budget for the shadow-ray wavefront (directions + visibility of one chunk of sample pairs).  B200 has 180 GB:
# 24 GB holds 93 of the 256 sample pairs of the 8 x 1024^2, n=16 workload per chunk.
SHADOW_SCRATCH_BUDGET = 24 << 30
_scratch_cache = {}


def _shadow_scratch(B, H, W, n_covered, n, dev):
    """Scratch for one traced env_shade call.  The library derives the chunking from the size it is given, so the buffer only
    ever grows, and with headroom: the covered-pixel count drifts from step to step, and re-allocating a multi-GB buffer on
    every small increase costs a cudaFree + cudaMalloc (hundreds of ms at 20 GB) per step."""
    nbytes = int(_lib.lib.gsb_env_shade_scratch_bytes(B, H, W, n_covered, n, SHADOW_SCRATCH_BUDGET))
    key = str(dev)
    buf = _scratch_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        want = min(max(nbytes, int(SHADOW_SCRATCH_BUDGET)), nbytes + nbytes // 4) if nbytes < SHADOW_SCRATCH_BUDGET else nbytes
        want = max(want, nbytes)
        _scratch_cache.pop(key, None)
        buf = None
        buf = _scratch_cache[key] = torch.empty(want, dtype=torch.uint8, device=dev)
    return buf


def _covered_pixels(mask):
    """Pixels that can emit shadow rays (one host read): the ray list of a chunk is sized for exactly these.  An understated
    count would lose rays, which the library reports as an error (gsb_env_shade_dropped_rays)."""
    return max(1, int(torch.count_nonzero(mask > 0)))


def _shade_kernels(B, H, W, n_covered, n, scratch):
    """Kernels one env_shade call launches: 1, or (generate + trace + shade) per chunk of sample pairs when tracing."""
    if scratch is None:
        return 1
    return 3 * int(_lib.lib.gsb_env_shade_chunks(B, H, W, n_covered, n, scratch.numel()))


def _cdf_top_tables(rows, cols):
    """Every 16th CDF entry (index 15, 31, ...) padded to 16 columns with 2.0: the top level of the kernel's 16-ary
    CDF search.  Two tiny torch ops per call on the 256x256 probe."""
    def top(c):
        t = c[..., 15::16]
        pad = 16 - t.shape[-1]
        if pad < 0 or c.shape[-1] % 16 != 0:
            return torch.full(c.shape[:-1] + (16,), 2.0, dtype=torch.float32, device=c.device)
        return torch.nn.functional.pad(t, (0, pad), value=2.0).contiguous()
    return top(rows), top(cols)


# Multi-GPU: views shard across ranks and cover different fractions of their frames, so the shadow rays -- 3/4 of the step -- are
# unevenly spread (8 ranks: 151-185 ms of tracing per rank, profiles/r2n).  With BALANCE_SHADING the forward pass deals the
# rows of every view out in blocks of 8 to all ranks (one all-to-all of the G-buffer, 68 B per pixel), each rank shades and traces
# its share of EVERY view, and a second all-to-all brings radiance and visibility bits home (24 + 2 n^2 / 8 B per pixel).  The
# backward pass stays local: it replays the visibility bits with the ids and the seed the forward used.
BALANCE_SHADING = True
PIXEL_ID_BASE = None    # tests: seed pixel i's sample stream with PIXEL_ID_BASE + i in the local path too (what the dealt path does with rank * B*H*W)
_ROW_BLOCK = 8          # rows per dealt block = the CTA tile height of the shading kernels


def _balance_world(H, B=None, device=None):
    """Ranks the forward shading pass is dealt out to: the world size when a process group is up, the rows split into whole
    8-row blocks and -- checked with one tiny all-gather when `B` is given -- every rank shades the same number of views
    (the exchange is an equal-split all-to-all); else 1 = every pixel is shaded at home.  All ranks reach the same answer."""
    if not BALANCE_SHADING:
        return 1
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 1
    world = dist.get_world_size()
    if world <= 1:
        return 1
    if B is not None:
        shapes = torch.empty((world, 2), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(shapes, torch.tensor([[B, H]], dtype=torch.int64, device=device))
        if not bool((shapes == shapes[0]).all()):
            return 1
    return world if H % (_ROW_BLOCK * world) == 0 else 1


def _deal_rows(x, world):
    """[B,H,W,C] -> [world, B, H/world, W, C]: row block j of every image goes to part j mod world."""
    B, H, W, C = x.shape
    return x.view(B, H // (_ROW_BLOCK * world), world, _ROW_BLOCK, W, C).permute(2, 0, 1, 3, 4, 5).reshape(world, B, H // world, W, C)


def _collect_rows(y, H):
    """inverse of _deal_rows: [world, B, H/world, W, C] -> [B,H,W,C]"""
    world, B, _, W, C = y.shape
    return y.view(world, B, H // (_ROW_BLOCK * world), _ROW_BLOCK, W, C).permute(1, 2, 0, 3, 4, 5).reshape(B, H, W, C)


def _exchange_out(pack, world):
    """[B,H,W,C] on every rank -> [world * B, H/world, W, C]: this rank's share of the rows of EVERY rank's images (source-rank major)."""
    import torch.distributed as dist
    B, H, W, C = pack.shape
    send = _deal_rows(pack, world).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    return recv.view(world * B, H // world, W, C)


def _exchange_home(out, world, H):
    """inverse route of _exchange_out: [world * B, H/world, W, C] per rank -> [B,H,W,C] at the owners"""
    import torch.distributed as dist
    Bs, Hb, W, C = out.shape
    back = torch.empty_like(out)
    dist.all_to_all_single(back, out.contiguous())
    return _collect_rows(back.view(world, Bs // world, Hb, W, C), H)


class _EnvShade(torch.autograd.Function):
    _random_perm = {}

    @staticmethod
    def _forward_balanced(optix_ctx, tens, seed, BSDF, n, shadow_scale, need_vis):
        """Forward pass with the pixels dealt out over the ranks (see BALANCE_SHADING).  -> diff, spec, vis bits, pixel ids, seed;
        all in the caller's own [B,H,W] layout."""
        import torch.distributed as dist
        mask, ro, pos, nrm, vpos, kd, ks = tens[:7]
        B, H, W = mask.shape
        dev = mask.device
        world, rank = dist.get_world_size(), dist.get_rank()
        Hb = H // world
        npix = B * H * W
        ids = torch.arange(rank * npix, (rank + 1) * npix, dtype=torch.int32, device=dev)
        pack = torch.cat([mask.unsqueeze(-1), ro, pos, nrm, kd, ks, ids.view(torch.float32).view(B, H, W, 1)], -1)      # 17 floats
        g = _exchange_out(pack, world)
        # per-view camera positions of every rank, and ONE seed for all (rank 0's): hash(seed, id) must agree between the rank
        # that shades a pixel and the rank that owns it
        head = torch.cat([vpos, torch.full((B, 1), seed & 0x7FFFFFFF, dtype=torch.int32, device=dev).view(torch.float32)], -1)
        heads = torch.empty((world * B, 4), dtype=torch.float32, device=dev)      # rank-major concatenation (the layout every backend accepts)
        dist.all_gather_into_tensor(heads, head.contiguous())
        heads = heads.view(world, B, 4)
        seed = int(heads[0, 0, 3:4].view(torch.int32).item())
        t = [g[..., 0].contiguous(), g[..., 1:4].contiguous(), g[..., 4:7].contiguous(), g[..., 7:10].contiguous(),
             heads[..., 0:3].reshape(world * B, 3).contiguous(), g[..., 10:13].contiguous(), g[..., 13:16].contiguous()] + list(tens[7:])
        gids = g[..., 16].contiguous().view(torch.int32)
        Bs = world * B
        n_cov = _covered_pixels(t[0])
        scratch = _shadow_scratch(Bs, Hb, W, n_cov, n, dev)
        words = (2 * n * n + 31) // 32
        out = torch.empty((Bs, Hb, W, 6 + words), dtype=torch.float32, device=dev)
        d_s = torch.empty((Bs, Hb, W, 3), dtype=torch.float32, device=dev)
        s_s = torch.empty((Bs, Hb, W, 3), dtype=torch.float32, device=dev)
        v_s = torch.empty((Bs * Hb * W, words), dtype=torch.int32, device=dev)
        ptrs = [_lib.ptr(x) for x in t]
        dims = (Bs, Hb, W, t[7].shape[0], t[7].shape[1], t[13].shape[0])
        _lib.check(_lib.lib.gsb_env_shade_fwd(*ptrs, *dims, BSDF, n, seed & 0xFFFFFFFF, shadow_scale, optix_ctx.bvh_ptr(), _lib.ptr(scratch),
                                              scratch.numel(), n_cov, _lib.ptr(v_s), _lib.ptr(d_s), _lib.ptr(s_s), _lib.ptr(gids),
                                              _lib.current_stream(dev)), "gsb_env_shade_fwd", kernels=_shade_kernels(Bs, Hb, W, n_cov, n, scratch))
        out[..., 0:3] = d_s
        out[..., 3:6] = s_s
        out[..., 6:] = v_s.view(torch.float32).view(Bs, Hb, W, words)
        home = _exchange_home(out, world, H)
        diff, spec = home[..., 0:3].contiguous(), home[..., 3:6].contiguous()
        vis = home[..., 6:].contiguous().view(torch.int32).view(npix, words) if need_vis else None
        return diff, spec, vis, ids, seed

    @staticmethod
    def perms(n, device):
        key = (n, str(device))
        if key not in _EnvShade._random_perm:
            # 32k random permutations decorrelating the light / BSDF strata (reference ops.py:87-89).  Drawn from a generator of
            # its own: the table is the same on every rank (pixels shaded away from home replay the same strata in the backward
            # pass at home) and building it does not advance the global RNG on the first call only
            gen = torch.Generator(device=device).manual_seed(0x9E3779B9 + n)
            _EnvShade._random_perm[key] = torch.argsort(torch.rand(32768, n * n, device=device, generator=gen), dim=-1).int().contiguous()
        return _EnvShade._random_perm[key]

    @staticmethod
    def _launch_args(t):
        mask, ro, pos, nrm, vpos, kd, ks, light, pdf, rows, cols, rows_top, cols_top, perms = t
        B, H, W = mask.shape
        return [_lib.ptr(x) for x in t], (B, H, W, light.shape[0], light.shape[1], perms.shape[0])

    @staticmethod
    def forward(ctx, optix_ctx, mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols, BSDF,
                n_samples_x, rnd_seed, shadow_scale, perms):
        seed = np.random.randint(2 ** 31) if rnd_seed is None else int(rnd_seed)
        B, H, W = mask.shape
        dev = gb_pos.device

        def dense(t, shape=None):
            t = t.detach().float()
            if shape is not None:
                t = t.expand(shape)
            return t.contiguous()
        full = (B, H, W, 3)
        tens = [dense(mask), dense(ro, full), dense(gb_pos, full), dense(gb_normal, full),
                dense(gb_view_pos.reshape(-1, 3).expand(B, 3)), dense(gb_kd, full), dense(gb_ks, full),
                dense(light), dense(pdf), dense(rows), dense(cols)]
        tens += list(_cdf_top_tables(tens[9], tens[10])) + [perms]
        ptrs, dims = _EnvShade._launch_args(tens)
        bvh = optix_ctx.bvh_ptr() if optix_ctx is not None else None
        # Visibility bits of every shadow ray are kept for the backward pass (which replays the same samples) unless the
        # caller asked for decorrelated fwd/bwd seeds: 2 n^2 bits per pixel instead of 2 n^2 more rays per pixel.
        vis = None
        scratch = None
        ids = None
        tracing = bvh is not None and float(shadow_scale) > 0
        n_cov = 0
        if tracing and rnd_seed is not None and _balance_world(H, B, dev) > 1:
            diff, spec, vis, ids, seed = _EnvShade._forward_balanced(optix_ctx, tens, seed, BSDF, n_samples_x, float(shadow_scale),
                                                                     any(ctx.needs_input_grad))
        else:
            diff = torch.empty(full, dtype=torch.float32, device=dev)
            spec = torch.empty(full, dtype=torch.float32, device=dev)
            if PIXEL_ID_BASE is not None:
                ids = torch.arange(PIXEL_ID_BASE, PIXEL_ID_BASE + B * H * W, dtype=torch.int32, device=dev)
            if tracing:
                # one host read: the ray list of a chunk is sized for the pixels that can emit rays, so views that cover 15 % of
                # the frame run in a quarter of the chunks (each chunk is three launches over all pixels)
                n_cov = _covered_pixels(tens[0])
                scratch = _shadow_scratch(B, H, W, n_cov, n_samples_x, dev)
                # (autograd.Function.forward runs under no_grad: ask the ctx whether a backward pass can follow)
                if rnd_seed is not None and any(ctx.needs_input_grad):
                    vis = torch.empty((B * H * W, (2 * n_samples_x * n_samples_x + 31) // 32), dtype=torch.int32, device=dev)
            _lib.check(_lib.lib.gsb_env_shade_fwd(*ptrs, *dims, BSDF, n_samples_x, seed & 0xFFFFFFFF, float(shadow_scale),
                                                  bvh, _lib.ptr(scratch), 0 if scratch is None else scratch.numel(), n_cov,
                                                  _lib.ptr(vis), _lib.ptr(diff), _lib.ptr(spec), _lib.ptr(ids), _lib.current_stream(dev)),
                       "gsb_env_shade_fwd", kernels=_shade_kernels(B, H, W, n_cov, n_samples_x, scratch))
        ctx.ids = ids
        ctx.seed = seed
        ctx.n_cov = n_cov
        ctx.vis = vis
        ctx.save_for_backward(*tens)
        ctx.optix_ctx = optix_ctx
        ctx.occluder_keep = None if optix_ctx is None else (optix_ctx.occluder, optix_ctx._keep)
        ctx.meta = (BSDF, n_samples_x, rnd_seed, float(shadow_scale), light.shape)
        return diff, spec

    @staticmethod
    def backward(ctx, g_diff, g_spec):
        tens = ctx.saved_tensors
        BSDF, n, rnd_seed, shadow_scale, light_shape = ctx.meta
        # decorrelated mode draws a fresh seed for the backward pass (reference ops.py:103)
        seed = np.random.randint(2 ** 31) if rnd_seed is None else int(ctx.seed)
        ptrs, dims = _EnvShade._launch_args(tens)
        dev = tens[2].device
        full = tens[2].shape
        gd, gs = g_diff.float().contiguous(), g_spec.float().contiguous()
        g_pos, g_nrm, g_kd, g_ks = (torch.empty(full, dtype=torch.float32, device=dev) for _ in range(4))
        g_light = torch.empty(tuple(light_shape[:2]) + (4,), dtype=torch.float32, device=dev)       # padded texels: vector reductions
        bvh = None
        if ctx.occluder_keep is not None and ctx.occluder_keep[0] is not None:
            bvh = ctx.occluder_keep[0].data_ptr()          # the occluder the forward pass traced against
        scratch = None
        if bvh is not None and shadow_scale > 0 and ctx.vis is None:      # decorrelated seeds: trace again
            B_, H_, W_ = dims[0], dims[1], dims[2]
            scratch = _shadow_scratch(B_, H_, W_, ctx.n_cov, n, dev)
        _lib.check(_lib.lib.gsb_env_shade_bwd(*ptrs, *dims, BSDF, n, seed & 0xFFFFFFFF, shadow_scale, bvh, _lib.ptr(scratch),
                                              0 if scratch is None else scratch.numel(), ctx.n_cov, _lib.ptr(ctx.vis),
                                              _lib.ptr(gd), _lib.ptr(gs), _lib.ptr(g_pos), _lib.ptr(g_nrm),
                                              _lib.ptr(g_kd), _lib.ptr(g_ks), _lib.ptr(g_light), _lib.ptr(ctx.ids),
                                              _lib.current_stream(dev)), "gsb_env_shade_bwd",
                   kernels=_shade_kernels(dims[0], dims[1], dims[2], ctx.n_cov, n, scratch))
        # same gradient set as the reference (ops.py:108): pos, normal, kd, ks, light
        return (None, None, None, g_pos, g_nrm, None, g_kd, g_ks, g_light[..., 0:3], None, None, None, None, None, None, None, None)


def optix_env_shade(optix_ctx, mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols,
                    BSDF="pbr", n_samples_x=8, rnd_seed=None, shadow_scale=1.0, perms=None):
    """Reference ops.py:141-143.  Returns (diffuse_accum, specular_accum) [B,H,W,3] (diffuse demodulated).
    `perms` (int32 [P, n^2]) overrides the cached permutation table (tests pass the oracle's table)."""
    _lib.require_cuda(gb_pos, "optix_env_shade")
    iBSDF = _BSDF.index(BSDF)
    if perms is None:
        perms = _EnvShade.perms(n_samples_x, gb_pos.device)
    return _EnvShade.apply(optix_ctx, mask, ro, gb_pos, gb_normal, gb_view_pos, gb_kd, gb_ks, light, pdf, rows, cols,
                           iBSDF, n_samples_x, rnd_seed, shadow_scale, perms.int().contiguous())


# ------------------------------------------------------------------------------------------------
def _view(t):
    """[B,H,W,C] tensor or channel-slice of one -> (tensor to keep alive, pixel stride) with dense rows/batches."""
    B, H, W, C = t.shape
    ps = t.stride(2)
    if t.stride(3) == 1 and t.stride(1) == W * ps and t.stride(0) == H * W * ps and t.dtype == torch.float32:
        return t, ps
    t = t.float().contiguous()
    return t, C


class _Bilateral(torch.autograd.Function):
    @staticmethod
    def forward(ctx, col_a, col_b, nrm, zdz, sigma):
        B, H, W, _ = col_a.shape
        dev = col_a.device
        ca, ps_c = _view(col_a.detach())
        cb = None
        if col_b is not None:
            cb, ps_b = _view(col_b.detach())
            if ps_b != ps_c:
                ca, cb, ps_c = ca.contiguous(), cb.contiguous(), 3
        n, ps_n = _view(nrm.detach())
        z, ps_z = _view(zdz.detach())
        out_a = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
        w_a = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        out_b = torch.empty_like(out_a) if cb is not None else None
        w_b = torch.empty_like(w_a) if cb is not None else None
        _lib.check(_lib.lib.gsb_bilateral_fwd(ca.data_ptr(), None if cb is None else cb.data_ptr(), n.data_ptr(),
                                              z.data_ptr(), ps_c, ps_n, ps_z, B, H, W, float(sigma), _lib.ptr(out_a),
                                              _lib.ptr(w_a), _lib.ptr(out_b), _lib.ptr(w_b), _lib.current_stream(dev)),
                   "gsb_bilateral_fwd")
        ctx.save_for_backward(n, z, w_a)
        ctx.meta = (ps_n, ps_z, float(sigma), cb is not None)
        if cb is None:
            return out_a
        return out_a, out_b

    @staticmethod
    def backward(ctx, g_a, g_b=None):
        n, z, w = ctx.saved_tensors
        ps_n, ps_z, sigma, pair = ctx.meta
        B, H, W = w.shape
        dev = w.device
        ga = g_a.float().contiguous()
        gb = g_b.float().contiguous() if pair else None
        d_a = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
        d_b = torch.empty_like(d_a) if pair else None
        scratch = torch.empty((2, B, H, W), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib.gsb_bilateral_bwd(_lib.ptr(ga), _lib.ptr(w), _lib.ptr(gb), _lib.ptr(w) if pair else None,
                                              n.data_ptr(), z.data_ptr(), ps_n, ps_z, B, H, W, sigma, _lib.ptr(d_a),
                                              _lib.ptr(d_b), _lib.ptr(scratch), _lib.current_stream(dev)),
                   "gsb_bilateral_bwd")
        return d_a, d_b, None, None, None


def bilateral_denoiser(col, nrm, zdz, sigma):
    """Reference ops.py:145-147: cross-bilateral filter, returns rgb / w.  Gradient flows to `col` only."""
    return _Bilateral.apply(col, None, nrm, zdz, sigma)


def bilateral_denoiser_pair(col_a, col_b, nrm, zdz, sigma):
    """Both images filtered in ONE pass with shared guide weights (the renderer denoises diffuse and specular
    light with identical guides, reference render.py:140-142)."""
    return _Bilateral.apply(col_a, col_b, nrm, zdz, sigma)
