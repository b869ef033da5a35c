// Fused G-buffer build and fused shade-combine + composite, forward and adjoint.
//
// The reference assembles these per iteration from ~120 PyTorch / nvdiffrast calls:
//   render_layer  render/render.py:199-317   five dr.interpolate calls (position, face normal, vertex normal, clip position with
//                                             screen derivatives, mSDF), face normals by gather/cross/normalise, z and |dz|
//   shade         render/render.py:55-63,100-118,144-186   jittered taps for the normal / material regularisers, demodulated
//                                             combine diff * kd * (1 - metal) + spec, eleven `cat`s
//   render_mesh   render/render.py:352-433   `lerp` composite of every buffer over the background
// Here: gsb_gbuffer_fwd/bwd = ONE pass over the pixels reading the rasteriser output once (barycentrics, triangle id) and
// gathering the three vertices once for every attribute; gsb_compose_fwd/bwd = ONE pass producing all composited buffers.
// Both are HBM-bound streaming kernels: algorithmic bytes = inputs read once + outputs written once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"
#include "vec.cuh"

using namespace gsb;

namespace {
constexpr int kThreads = 256;
inline int nblk(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }
constexpr float kDepthEps = 0.00001f;

__device__ __forceinline__ V3 ldv(const float* p, int64_t i) { return v3(__ldg(p + i * 3), __ldg(p + i * 3 + 1), __ldg(p + i * 3 + 2)); }
__device__ __forceinline__ void stv(float* p, int64_t i, V3 a) { p[i * 3] = a.x; p[i * 3 + 1] = a.y; p[i * 3 + 2] = a.z; }
__device__ __forceinline__ void addv(float* p, int64_t i, V3 a) {
  atomicAdd(p + i * 3, a.x); atomicAdd(p + i * 3 + 1, a.y); atomicAdd(p + i * 3 + 2, a.z);
}

struct GBufArgs {
  const float4 *rast, *rast_db;      // [n_pix] (u, v, z/w, id+1), (du/dX, du/dY, dv/dX, dv/dY)
  const float *v_pos, *v_nrm, *msdf; // [V,3], [V,3], [V] (msdf may be null)
  const float4* v_clip;              // [B,V] clip-space positions
  const int32_t* tris;               // [F,3]
  int64_t n_pix, pix_per_img;
  int n_verts;
};

__global__ void __launch_bounds__(kThreads) k_gbuffer_fwd(GBufArgs a, float* __restrict__ pos, float* __restrict__ nrm,
                                                          float* __restrict__ geo, float* __restrict__ depth, float* __restrict__ msdf_img) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n_pix) return;
  const float4 r = __ldg(a.rast + i);
  const int f = (int)r.w - 1;
  if (f < 0) {
    stv(pos, i, v3(0.f)); stv(nrm, i, v3(0.f)); stv(geo, i, v3(0.f));
    depth[i * 2] = 1.f; depth[i * 2 + 1] = 0.f;               // clamp(0, eps) / clamp(0, eps), as the reference computes it
    if (msdf_img) msdf_img[i] = 0.f;
    return;
  }
  const int i0 = __ldg(a.tris + (size_t)f * 3), i1 = __ldg(a.tris + (size_t)f * 3 + 1), i2 = __ldg(a.tris + (size_t)f * 3 + 2);
  const float u = r.x, v = r.y, w = 1.f - r.x - r.y;
  const V3 p0 = ldv(a.v_pos, i0), p1 = ldv(a.v_pos, i1), p2 = ldv(a.v_pos, i2);
  stv(pos, i, p0 * u + p1 * v + p2 * w);
  stv(nrm, i, ldv(a.v_nrm, i0) * u + ldv(a.v_nrm, i1) * v + ldv(a.v_nrm, i2) * w);
  const V3 c = cross(p1 - p0, p2 - p0);
  stv(geo, i, c * rsqrtf(fmaxf(dot(c, c), 1e-20f)));
  const size_t off = (size_t)(i / a.pix_per_img) * a.n_verts;
  const float4 c0 = __ldg(a.v_clip + off + i0), c1 = __ldg(a.v_clip + off + i1), c2 = __ldg(a.v_clip + off + i2);
  const float4 d = __ldg(a.rast_db + i);
  const float cz = u * c0.z + v * c1.z + w * c2.z, cw = u * c0.w + v * c1.w + w * c2.w;
  // the reference reads channels [2] and [3] of the interleaved (d/dX, d/dY) derivative tensor, i.e. the screen derivatives of
  // clip.y (render.py:280-285); reproduced as written
  const float dyx = d.x * (c0.y - c2.y) + d.z * (c1.y - c2.y), dyy = d.y * (c0.y - c2.y) + d.w * (c1.y - c2.y);
  const float z0 = fmaxf(cz, kDepthEps) / fmaxf(cw, kDepthEps);
  const float z1 = fmaxf(cz + fabsf(dyx), kDepthEps) / fmaxf(cw + fabsf(dyy), kDepthEps);
  depth[i * 2] = z0;
  depth[i * 2 + 1] = fabsf(z1 - z0);
  if (msdf_img) msdf_img[i] = u * __ldg(a.msdf + i0) + v * __ldg(a.msdf + i1) + w * __ldg(a.msdf + i2);
}

// g_v_pos, g_v_nrm, g_msdf: accumulated with atomics (zeroed by the caller); g_rast fully written
__global__ void __launch_bounds__(kThreads) k_gbuffer_bwd(GBufArgs a, const float* __restrict__ g_pos, const float* __restrict__ g_nrm,
                                                          const float* __restrict__ g_geo, const float* __restrict__ g_msdf_img,
                                                          float* __restrict__ g_v_pos, float* __restrict__ g_v_nrm, float* __restrict__ g_msdf,
                                                          float4* __restrict__ g_rast) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n_pix) return;
  const float4 r = __ldg(a.rast + i);
  const int f = (int)r.w - 1;
  if (f < 0) {
    if (g_rast) g_rast[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int i0 = __ldg(a.tris + (size_t)f * 3), i1 = __ldg(a.tris + (size_t)f * 3 + 1), i2 = __ldg(a.tris + (size_t)f * 3 + 2);
  const float u = r.x, v = r.y, w = 1.f - r.x - r.y;
  float gu = 0.f, gv = 0.f;
  const V3 p0 = ldv(a.v_pos, i0), p1 = ldv(a.v_pos, i1), p2 = ldv(a.v_pos, i2);
  V3 a0 = v3(0.f), a1 = v3(0.f), a2 = v3(0.f);           // gradient w.r.t. the three vertex positions
  if (g_pos) {
    const V3 g = ldv(g_pos, i);
    a0 = g * u; a1 = g * v; a2 = g * w;
    gu += dot(g, p0 - p2);
    gv += dot(g, p1 - p2);
  }
  if (g_geo) {
    const V3 g = ldv(g_geo, i), e1 = p1 - p0, e2 = p2 - p0, c = cross(e1, e2);
    const float l2 = dot(c, c);
    if (l2 > 1e-20f) {
      const float il = rsqrtf(l2);
      const V3 n = c * il, gc = (g - n * dot(n, g)) * il;
      const V3 ge1 = cross(e2, gc), ge2 = cross(gc, e1);
      a1 += ge1; a2 += ge2; a0 -= ge1 + ge2;
    }
  }
  if (g_v_pos && (g_pos || g_geo)) { addv(g_v_pos, i0, a0); addv(g_v_pos, i1, a1); addv(g_v_pos, i2, a2); }
  if (g_nrm) {
    const V3 g = ldv(g_nrm, i);
    if (g_v_nrm) { addv(g_v_nrm, i0, g * u); addv(g_v_nrm, i1, g * v); addv(g_v_nrm, i2, g * w); }
    const V3 n2 = ldv(a.v_nrm, i2);
    gu += dot(g, ldv(a.v_nrm, i0) - n2);
    gv += dot(g, ldv(a.v_nrm, i1) - n2);
  }
  if (g_msdf_img && a.msdf) {
    const float g = __ldg(g_msdf_img + i);
    if (g_msdf && g != 0.f) { atomicAdd(g_msdf + i0, g * u); atomicAdd(g_msdf + i1, g * v); atomicAdd(g_msdf + i2, g * w); }
    const float m2 = __ldg(a.msdf + i2);
    gu += g * (__ldg(a.msdf + i0) - m2);
    gv += g * (__ldg(a.msdf + i1) - m2);
  }
  if (g_rast) g_rast[i] = make_float4(gu, gv, 0.f, 0.f);
}

// ---- shade combine + composite --------------------------------------------------------------------------------------------
struct ComposeArgs {
  const float4* rast;                     // coverage = id+1 > 0
  const float *jitter;                    // [n_pix,2] uv offset of the regulariser tap (normal(0, 0.005)), or null (no taps)
  const float *gb_nrm;                    // [n_pix,3] interpolated vertex normal (before bending): normal-smoothness tap
  const float *tex, *tex_j;               // [n_pix,6] (kd, ks) and the same field sampled at a jittered position
  const float *sh_nrm, *geo_nrm, *depth;  // [n_pix,3], [n_pix,3], [n_pix,2]
  const float *diff, *spec;               // [n_pix,3] light accumulators (null: no 'diffuse_light' / 'specular_light' buffers)
  const float *col;                       // [n_pix,3] colour override (mode 2)
  const float *msdf_img;                  // [n_pix] or null
  const float *bg;                        // [n_pix,3] or [pix_per_img,3] (bg_batched = 0) or null (black)
  int64_t n_pix, pix_per_img;
  int composite;                          // 1: outputs are laid over the background with the coverage as alpha; 0: alpha = 1 (shade()'s own buffers)
  int H, W, mode, bg_batched;             // mode 0 'pbr': diff kd (1 - metal) + spec;  1 'diffuse': diff kd;  2: col;  3 'white': diff
};
struct ComposeOut {
  float *shaded, *z_grad, *normal, *geometric_normal, *kd, *ks, *kd_grad, *ks_grad, *normal_grad, *diffuse_light, *specular_light;   // [n_pix,4]
  float *msdf_image;                                                                                                               // [n_pix]
};

__device__ __forceinline__ void st4(float* p, int64_t i, float x, float y, float z, float w) {
  if (p) reinterpret_cast<float4*>(p)[i] = make_float4(x, y, z, w);
}

// bilinear tap positions / weights of F.grid_sample(align_corners=False, padding_mode='border') at pixel centre + offset
struct Tap { int x0, x1, y0, y1; float wx, wy; };
__device__ __forceinline__ Tap make_tap(int px, int py, float ox, float oy, int W, int H) {
  float x = fminf(fmaxf((float)px + ox * (float)W, 0.f), (float)(W - 1));
  float y = fminf(fmaxf((float)py + oy * (float)H, 0.f), (float)(H - 1));
  Tap t;
  t.x0 = (int)floorf(x); t.y0 = (int)floorf(y);
  t.wx = x - (float)t.x0; t.wy = y - (float)t.y0;
  t.x1 = min(t.x0 + 1, W - 1); t.y1 = min(t.y0 + 1, H - 1);
  return t;
}

__global__ void __launch_bounds__(kThreads) k_compose_fwd(ComposeArgs a, ComposeOut o) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n_pix) return;
  const float cov = __ldg(reinterpret_cast<const float*>(a.rast + i) + 3) > 0.f ? 1.f : 0.f;
  const float al = a.composite ? cov : 1.f;
  const V3 kd = ldv(a.tex, i * 2), ks = ldv(a.tex, i * 2 + 1);
  const V3 kj = ldv(a.tex_j, i * 2), sj = ldv(a.tex_j, i * 2 + 1);
  V3 col;
  V3 df = v3(0.f), sp = v3(0.f);
  if (a.diff) { df = ldv(a.diff, i); sp = ldv(a.spec, i); }
  // the 'kd' buffer holds what multiplied the diffuse light: the reference re-binds kd before building it (render.py:150-158)
  const V3 kd_eff = a.mode == 0 ? kd * (1.f - ks.z) : (a.mode == 3 ? v3(1.f) : kd);
  if (a.mode == 0) col = df * kd_eff + sp;
  else if (a.mode == 1 || a.mode == 3) col = df * kd_eff;
  else col = ldv(a.col, i);
  V3 bg = v3(0.f);
  if (a.bg) bg = ldv(a.bg, a.bg_batched ? i : i % a.pix_per_img);
  st4(o.shaded, i, bg.x + al * (col.x - bg.x), bg.y + al * (col.y - bg.y), bg.z + al * (col.z - bg.z), al);
  st4(o.kd, i, al * kd_eff.x, al * kd_eff.y, al * kd_eff.z, al);
  st4(o.ks, i, al * ks.x, al * ks.y, al * ks.z, al);
  st4(o.kd_grad, i, al * fabsf(kj.x - kd.x), al * fabsf(kj.y - kd.y), al * fabsf(kj.z - kd.z), al);
  st4(o.ks_grad, i, 0.f, al * fabsf(sj.y - ks.y), al * fabsf(sj.z - ks.z), al);
  if (o.normal_grad) {
    V3 ng = v3(0.f);
    if (a.jitter && cov > 0.f) {
      const int px = (int)(i % a.W), py = (int)((i / a.W) % a.H);
      const int64_t base = i - (int64_t)py * a.W - px;
      const Tap t = make_tap(px, py, __ldg(a.jitter + i * 2), __ldg(a.jitter + i * 2 + 1), a.W, a.H);
      const int64_t q00 = base + (int64_t)t.y0 * a.W + t.x0, q01 = base + (int64_t)t.y0 * a.W + t.x1, q10 = base + (int64_t)t.y1 * a.W + t.x0,
                    q11 = base + (int64_t)t.y1 * a.W + t.x1;
      const float w00 = (1.f - t.wx) * (1.f - t.wy), w01 = t.wx * (1.f - t.wy), w10 = (1.f - t.wx) * t.wy, w11 = t.wx * t.wy;
      auto cov = [&](int64_t q) { return __ldg(reinterpret_cast<const float*>(a.rast + q) + 3) > 0.f ? 1.f : 0.f; };
      const float mask_tap = w00 * cov(q00) + w01 * cov(q01) + w10 * cov(q10) + w11 * cov(q11);
      const V3 nj = ldv(a.gb_nrm, q00) * w00 + ldv(a.gb_nrm, q01) * w01 + ldv(a.gb_nrm, q10) * w10 + ldv(a.gb_nrm, q11) * w11;
      const V3 n = ldv(a.gb_nrm, i);
      ng = v3(fabsf(nj.x - n.x), fabsf(nj.y - n.y), fabsf(nj.z - n.z)) * mask_tap;       // grad_weight = mask * mask_tap, mask = 1 here
    }
    st4(o.normal_grad, i, ng.x, ng.y, ng.z, al);
  }
  if (o.normal) { const V3 n = ldv(a.sh_nrm, i); st4(o.normal, i, al * n.x, al * n.y, al * n.z, al); }
  if (o.geometric_normal) { const V3 n = ldv(a.geo_nrm, i); st4(o.geometric_normal, i, al * n.x, al * n.y, al * n.z, al); }
  if (o.z_grad) st4(o.z_grad, i, al * __ldg(a.depth + i * 2), al * __ldg(a.depth + i * 2 + 1), 0.f, al);
  if (a.diff) {
    st4(o.diffuse_light, i, al * df.x, al * df.y, al * df.z, al);
    st4(o.specular_light, i, al * sp.x, al * sp.y, al * sp.z, al);
  }
  if (o.msdf_image && a.msdf_img) {
    // a one-channel buffer has no alpha of its own: the reference's composite then takes the VALUE as alpha (render.py:352-357),
    // lerp(0, 1, coverage * msdf) = coverage * msdf, still one channel
    o.msdf_image[i] = al * __ldg(a.msdf_img + i);
  }
}

struct ComposeGradIn { const float *shaded, *kd, *ks, *kd_grad, *ks_grad, *normal_grad, *diffuse_light, *specular_light, *msdf_image; };   // null = no gradient
struct ComposeGradOut { float *diff, *spec, *col, *tex, *tex_j, *gb_nrm, *msdf_img; };                                                   // null = not wanted

__device__ __forceinline__ V3 ld4v(const float* p, int64_t i) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
  return v3(t.x, t.y, t.z);
}
__device__ __forceinline__ V3 sgn3(V3 a) {
  return v3((float)((a.x > 0.f) - (a.x < 0.f)), (float)((a.y > 0.f) - (a.y < 0.f)), (float)((a.z > 0.f) - (a.z < 0.f)));
}

// g.gb_nrm is accumulated with atomics (taps scatter; zeroed by the caller), everything else is fully written
__global__ void __launch_bounds__(kThreads) k_compose_bwd(ComposeArgs a, ComposeGradIn gi, ComposeGradOut go) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= a.n_pix) return;
  const float cov = __ldg(reinterpret_cast<const float*>(a.rast + i) + 3) > 0.f ? 1.f : 0.f;
  const float al = a.composite ? cov : 1.f;
  V3 g_df = v3(0.f), g_sp = v3(0.f), g_col = v3(0.f), g_kd = v3(0.f), g_ks = v3(0.f), g_kj = v3(0.f), g_sj = v3(0.f);
  if (al > 0.f) {
    const V3 kd = ldv(a.tex, i * 2), ks = ldv(a.tex, i * 2 + 1);
    if (gi.shaded) {
      const V3 g = ld4v(gi.shaded, i);
      if (a.mode == 2) {
        g_col = g;
      } else if (a.mode == 3) {
        g_df += g;
      } else {
        const V3 df = ldv(a.diff, i);
        const float k = a.mode == 0 ? 1.f - ks.z : 1.f;
        g_df += g * kd * k;
        g_kd += g * df * k;
        if (a.mode == 0) {
          g_sp += g;
          g_ks.z -= dot(g, df * kd);
        }
      }
    }
    if (gi.kd && a.mode != 3) {
      const V3 g = ld4v(gi.kd, i);
      if (a.mode == 0) {
        g_kd += g * (1.f - ks.z);
        g_ks.z -= dot(g, kd);
      } else {
        g_kd += g;
      }
    }
    if (gi.ks) g_ks += ld4v(gi.ks, i);
    if (gi.kd_grad) {
      const V3 s = sgn3(ldv(a.tex_j, i * 2) - kd) * ld4v(gi.kd_grad, i);
      g_kj += s; g_kd -= s;
    }
    if (gi.ks_grad) {
      V3 s = sgn3(ldv(a.tex_j, i * 2 + 1) - ks) * ld4v(gi.ks_grad, i);
      s.x = 0.f;
      g_sj += s; g_ks -= s;
    }
    if (gi.diffuse_light) g_df += ld4v(gi.diffuse_light, i);
    if (gi.specular_light) g_sp += ld4v(gi.specular_light, i);
    if (gi.normal_grad && go.gb_nrm && a.jitter && cov > 0.f) {
      const int px = (int)(i % a.W), py = (int)((i / a.W) % a.H);
      const int64_t base = i - (int64_t)py * a.W - px;
      const Tap t = make_tap(px, py, __ldg(a.jitter + i * 2), __ldg(a.jitter + i * 2 + 1), a.W, a.H);
      const int64_t q00 = base + (int64_t)t.y0 * a.W + t.x0, q01 = base + (int64_t)t.y0 * a.W + t.x1, q10 = base + (int64_t)t.y1 * a.W + t.x0,
                    q11 = base + (int64_t)t.y1 * a.W + t.x1;
      const float w00 = (1.f - t.wx) * (1.f - t.wy), w01 = t.wx * (1.f - t.wy), w10 = (1.f - t.wx) * t.wy, w11 = t.wx * t.wy;
      auto cov = [&](int64_t q) { return __ldg(reinterpret_cast<const float*>(a.rast + q) + 3) > 0.f ? 1.f : 0.f; };
      const float mask_tap = w00 * cov(q00) + w01 * cov(q01) + w10 * cov(q10) + w11 * cov(q11);
      const V3 nj = ldv(a.gb_nrm, q00) * w00 + ldv(a.gb_nrm, q01) * w01 + ldv(a.gb_nrm, q10) * w10 + ldv(a.gb_nrm, q11) * w11;
      const V3 s = sgn3(nj - ldv(a.gb_nrm, i)) * ld4v(gi.normal_grad, i) * mask_tap;
      addv(go.gb_nrm, i, v3(0.f) - s);
      addv(go.gb_nrm, q00, s * w00); addv(go.gb_nrm, q01, s * w01); addv(go.gb_nrm, q10, s * w10); addv(go.gb_nrm, q11, s * w11);
    }
  }
  if (go.diff) stv(go.diff, i, g_df);
  if (go.spec) stv(go.spec, i, g_sp);
  if (go.col) stv(go.col, i, g_col);
  if (go.tex) { stv(go.tex, i * 2, g_kd); stv(go.tex, i * 2 + 1, g_ks); }
  if (go.tex_j) { stv(go.tex_j, i * 2, g_kj); stv(go.tex_j, i * 2 + 1, g_sj); }
  if (go.msdf_img) go.msdf_img[i] = (gi.msdf_image && al > 0.f) ? __ldg(gi.msdf_image + i) : 0.f;
}

}  // namespace

extern "C" {

int gsb_gbuffer_fwd(const float* rast, const float* rast_db, const float* v_pos, const float* v_nrm, const float* msdf, const float* v_clip,
                    const int32_t* tris, int64_t n_batch, int64_t H, int64_t W, int64_t n_verts, float* pos, float* nrm, float* geo_nrm,
                    float* depth, float* msdf_img, void* stream_) {
  GBufArgs a;
  a.rast = (const float4*)rast; a.rast_db = (const float4*)rast_db; a.v_pos = v_pos; a.v_nrm = v_nrm; a.msdf = msdf;
  a.v_clip = (const float4*)v_clip; a.tris = tris; a.n_pix = n_batch * H * W; a.pix_per_img = H * W; a.n_verts = (int)n_verts;
  if (a.n_pix == 0) return 0;
  k_gbuffer_fwd<<<nblk(a.n_pix), kThreads, 0, (cudaStream_t)stream_>>>(a, pos, nrm, geo_nrm, depth, msdf ? msdf_img : nullptr);
  return (int)cudaGetLastError();
}

int gsb_gbuffer_bwd(const float* rast, const float* v_pos, const float* v_nrm, const float* msdf, const int32_t* tris, int64_t n_batch,
                    int64_t H, int64_t W, int64_t n_verts, const float* g_pos, const float* g_nrm, const float* g_geo_nrm,
                    const float* g_msdf_img, float* g_v_pos, float* g_v_nrm, float* g_msdf, float* g_rast, void* stream_) {
  GBufArgs a;
  a.rast = (const float4*)rast; a.rast_db = nullptr; a.v_pos = v_pos; a.v_nrm = v_nrm; a.msdf = msdf; a.v_clip = nullptr; a.tris = tris;
  a.n_pix = n_batch * H * W; a.pix_per_img = H * W; a.n_verts = (int)n_verts;
  if (a.n_pix == 0) return 0;
  k_gbuffer_bwd<<<nblk(a.n_pix), kThreads, 0, (cudaStream_t)stream_>>>(a, g_pos, g_nrm, g_geo_nrm, g_msdf_img, g_v_pos, g_v_nrm, g_msdf,
                                                                      (float4*)g_rast);
  return (int)cudaGetLastError();
}

/* pointer tables: in12 = {rast, jitter, gb_nrm, tex, tex_j, sh_nrm, geo_nrm, depth, diff, spec, col, msdf_img} then bg;
 * out12 = {shaded, z_grad, normal, geometric_normal, kd, ks, kd_grad, ks_grad, normal_grad, diffuse_light, specular_light, msdf_image} */
static ComposeArgs compose_args(const void* const* in12, const float* bg, int bg_batched, int64_t n_batch, int64_t H, int64_t W, int mode,
                                int composite) {
  ComposeArgs a;
  a.rast = (const float4*)in12[0]; a.jitter = (const float*)in12[1]; a.gb_nrm = (const float*)in12[2]; a.tex = (const float*)in12[3];
  a.tex_j = (const float*)in12[4]; a.sh_nrm = (const float*)in12[5]; a.geo_nrm = (const float*)in12[6]; a.depth = (const float*)in12[7];
  a.diff = (const float*)in12[8]; a.spec = (const float*)in12[9]; a.col = (const float*)in12[10]; a.msdf_img = (const float*)in12[11];
  a.bg = bg; a.bg_batched = bg_batched; a.n_pix = n_batch * H * W; a.pix_per_img = H * W; a.H = (int)H; a.W = (int)W; a.mode = mode;
  a.composite = composite;
  return a;
}

int gsb_compose_fwd(const void* const* in12, const float* bg, int bg_batched, int64_t n_batch, int64_t H, int64_t W, int mode, int composite,
                    void* const* out12, void* stream_) {
  const ComposeArgs a = compose_args(in12, bg, bg_batched, n_batch, H, W, mode, composite);
  if (a.n_pix == 0) return 0;
  if (mode < 0 || mode > 3 || !a.rast || !a.tex || !a.tex_j || (mode != 2 && (!a.diff || !a.spec)) || (mode == 2 && !a.col))
    return (int)cudaErrorInvalidValue;
  ComposeOut o;
  float* const* p = (float* const*)out12;
  o.shaded = p[0]; o.z_grad = p[1]; o.normal = p[2]; o.geometric_normal = p[3]; o.kd = p[4]; o.ks = p[5]; o.kd_grad = p[6]; o.ks_grad = p[7];
  o.normal_grad = p[8]; o.diffuse_light = p[9]; o.specular_light = p[10]; o.msdf_image = p[11];
  k_compose_fwd<<<nblk(a.n_pix), kThreads, 0, (cudaStream_t)stream_>>>(a, o);
  return (int)cudaGetLastError();
}

/* gin9 = gradients of {shaded, kd, ks, kd_grad, ks_grad, normal_grad, diffuse_light, specular_light, msdf_image} (NULL = none);
 * gout7 = {diff, spec, col, tex, tex_j, gb_nrm (accumulated, zero it first), msdf_img} (NULL = not wanted) */
int gsb_compose_bwd(const void* const* in12, int64_t n_batch, int64_t H, int64_t W, int mode, int composite, const void* const* gin9,
                    void* const* gout7, void* stream_) {
  const ComposeArgs a = compose_args(in12, nullptr, 0, n_batch, H, W, mode, composite);
  if (a.n_pix == 0) return 0;
  ComposeGradIn gi;
  const float* const* q = (const float* const*)gin9;
  gi.shaded = q[0]; gi.kd = q[1]; gi.ks = q[2]; gi.kd_grad = q[3]; gi.ks_grad = q[4]; gi.normal_grad = q[5]; gi.diffuse_light = q[6];
  gi.specular_light = q[7]; gi.msdf_image = q[8];
  ComposeGradOut go;
  float* const* p = (float* const*)gout7;
  go.diff = p[0]; go.spec = p[1]; go.col = p[2]; go.tex = p[3]; go.tex_j = p[4]; go.gb_nrm = p[5]; go.msdf_img = p[6];
  k_compose_bwd<<<nblk(a.n_pix), kThreads, 0, (cudaStream_t)stream_>>>(a, gi, go);
  return (int)cudaGetLastError();
}

}  // extern "C"
