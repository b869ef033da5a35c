// HBM-bound per-vertex / per-pixel operators of the deferred renderer for sm_100a:
//   xfm_points              (reference render/renderutils/c_src/mesh.cu:22,56)
//   prepare_shading_normal  (reference render/renderutils/c_src/normal.cu:98,128)
//   image_loss              (reference render/renderutils/c_src/loss.cu:95,137)
// One thread per vertex / pixel, NHWC fp32, coalesced 12-byte rows; broadcast inputs via strides.
// No tensor cores (no contraction larger than 4x4), roofline = HBM.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"
#include "vec.cuh"

using namespace gsb;

namespace {

constexpr int kThreads = 256;

inline int blocks_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

// ---------------------------------------------------------------------------------------------
// xfm_points: out[b,n,:] = M_b * (p_n, 1)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_xfm_fwd(const float* __restrict__ pts, const float* __restrict__ mtx,
                                                      int64_t n_pts, int pts_batched, float* __restrict__ out) {
  __shared__ float m[16];
  const int b = blockIdx.y;
  if (threadIdx.x < 16) m[threadIdx.x] = mtx[b * 16 + threadIdx.x];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_pts) return;
  V3 p = ld3(pts + ((pts_batched ? (int64_t)b * n_pts : 0) + i) * 3);
  float4 o;
  o.x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
  o.y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
  o.z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
  o.w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
  reinterpret_cast<float4*>(out)[(int64_t)b * n_pts + i] = o;
}

// g_points[n] = sum_b M_b[:, :3]^T g_out[b, n]   (sum over b only when the points are shared)
__global__ void __launch_bounds__(kThreads) k_xfm_bwd(const float* __restrict__ mtx, const float* __restrict__ g_out,
                                                      int n_batch, int64_t n_pts, int pts_batched,
                                                      float* __restrict__ g_pts) {
  extern __shared__ float m[];  // [n_batch][16]
  for (int k = threadIdx.x; k < n_batch * 16; k += kThreads) m[k] = mtx[k];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_pts) return;
  const int b0 = pts_batched ? blockIdx.y : 0, b1 = pts_batched ? blockIdx.y + 1 : n_batch;
  V3 acc = v3(0.f);
  for (int b = b0; b < b1; ++b) {
    float4 g = __ldg(reinterpret_cast<const float4*>(g_out) + (int64_t)b * n_pts + i);
    const float* M = m + b * 16;
    acc.x += M[0] * g.x + M[4] * g.y + M[8] * g.z + M[12] * g.w;
    acc.y += M[1] * g.x + M[5] * g.y + M[9] * g.z + M[13] * g.w;
    acc.z += M[2] * g.x + M[6] * g.y + M[10] * g.z + M[14] * g.w;
  }
  st3(g_pts + ((pts_batched ? (int64_t)blockIdx.y * n_pts : 0) + i) * 3, acc);
}

// ---------------------------------------------------------------------------------------------
// prepare_shading_normal
// ---------------------------------------------------------------------------------------------
struct NormalArgs {
  const float* in[6];   // pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm
  int64_t sb[6], sy[6], sx[6];
  float* gout[6];       // backward: full-resolution [B,H,W,3] gradient per input (may be null)
  const float* g;       // backward: d/d out
  float* out;           // forward
  int B, H, W, two_sided, opengl;
};

__device__ __forceinline__ V3 fetch(const NormalArgs& a, int k, int b, int y, int x) {
  return ld3(a.in[k] + b * a.sb[k] + y * a.sy[k] + x * a.sx[k]);
}

constexpr float kNormalThreshold = 0.1f;

template <bool BWD>
__global__ void __launch_bounds__(kThreads) k_shading_normal(NormalArgs a) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int64_t npix = (int64_t)a.B * a.H * a.W;
  if (i >= npix) return;
  const int x = (int)(i % a.W), y = (int)((i / a.W) % a.H), b = (int)(i / ((int64_t)a.W * a.H));
  const V3 pos = fetch(a, 0, b, y, x), vpos = fetch(a, 1, b, y, x), pert = fetch(a, 2, b, y, x);
  const V3 sn_raw = fetch(a, 3, b, y, x), st_raw = fetch(a, 4, b, y, x), gn = fetch(a, 5, b, y, x);
  const float sgn = a.opengl ? -1.f : 1.f;

  const V3 sn = normalize0(sn_raw), st = normalize0(st_raw);
  const V3 view_raw = vpos - pos, view = normalize0(view_raw);
  const V3 bit_raw = cross(st, sn), bit = normalize0(bit_raw);
  const float pz = fmaxf(pert.z, 0.f);
  const V3 sh_raw = st * pert.x + bit * (sgn * pert.y) + sn * pz;
  const V3 sh = normalize0(sh_raw);
  const bool flip = a.two_sided && dot(view, gn) < 0.f;
  const V3 s_n = flip ? -sh : sh, g_n = flip ? -gn : gn;
  const float dp = dot(view, s_n);
  const float t = clampf(dp / kNormalThreshold, 0.f, 1.f);
  if (!BWD) {
    st3(a.out + i * 3, g_n * (1.f - t) + s_n * t);
    return;
  }
  // ---- adjoint (same case analysis as normal.cu:72-96,152-180) ----
  const V3 go = ld3(a.g + i * 3);
  V3 d_view = v3(0.f), d_sn2 = v3(0.f), d_gn2 = v3(0.f);   // w.r.t. (possibly flipped) s_n, g_n
  if (dp > kNormalThreshold) {
    d_sn2 = go;
  } else {
    d_gn2 = go * (1.f - t);
    d_sn2 = go * t;
    float d_t = dot(go, s_n - g_n);
    float d_dp = (dp < 0.f || dp > kNormalThreshold) ? 0.f : d_t / kNormalThreshold;
    d_view += s_n * d_dp;
    d_sn2 += view * d_dp;
  }
  V3 d_sh = flip ? -d_sn2 : d_sn2;
  V3 d_gn = flip ? -d_gn2 : d_gn2;
  // shading = normalize(st*px + sgn*bit*py + sn*max(pz,0))
  V3 d_sh_raw = normalize0_bwd(sh_raw, d_sh);
  V3 d_pert = v3(0.f), d_sn = v3(0.f), d_st = v3(0.f), d_bit = v3(0.f);
  if (pert.z > 0.f) {
    d_sn += d_sh_raw * pert.z;
    d_pert.z += dot(d_sh_raw, sn);
  }
  d_bit += d_sh_raw * (sgn * pert.y);
  d_pert.y += sgn * dot(d_sh_raw, bit);
  d_st += d_sh_raw * pert.x;
  d_pert.x += dot(d_sh_raw, st);
  V3 d_bit_raw = normalize0_bwd(bit_raw, d_bit);
  cross_bwd(st, sn, d_bit_raw, d_st, d_sn);
  V3 d_view_raw = normalize0_bwd(view_raw, d_view);
  V3 d_sn_raw = normalize0_bwd(sn_raw, d_sn);
  V3 d_st_raw = normalize0_bwd(st_raw, d_st);
  if (a.gout[0]) st3(a.gout[0] + i * 3, -d_view_raw);
  if (a.gout[1]) st3(a.gout[1] + i * 3, d_view_raw);
  if (a.gout[2]) st3(a.gout[2] + i * 3, d_pert);
  if (a.gout[3]) st3(a.gout[3] + i * 3, d_sn_raw);
  if (a.gout[4]) st3(a.gout[4] + i * 3, d_st_raw);
  if (a.gout[5]) st3(a.gout[5] + i * 3, d_gn);
}

// ---------------------------------------------------------------------------------------------
// image_loss
// ---------------------------------------------------------------------------------------------
enum { LOSS_L1 = 0, LOSS_MSE = 1, LOSS_RELMSE = 2, LOSS_SMAPE = 3 };
enum { TM_NONE = 0, TM_LOG_SRGB = 1 };

__device__ __forceinline__ float srgb_fwd(float x) {
  return x > 0.0031308f ? powf(fmaxf(x, 0.0031308f), 1.0f / 2.4f) * 1.055f - 0.055f : 12.92f * fmaxf(x, 0.f);
}
// d/dx of srgb_fwd (loss.cu:31-37)
__device__ __forceinline__ float srgb_bwd(float x) {
  if (x > 0.0031308f) return 0.439583f / powf(x, 0.583333f);
  return x > 0.f ? 12.92f : 0.f;
}
__device__ __forceinline__ float tonemap(float x, int tm) { return tm == TM_LOG_SRGB ? srgb_fwd(logf(x + 1.f)) : x; }
__device__ __forceinline__ float sign0(float x) { return x == 0.f ? 0.f : (x < 0.f ? -1.f : 1.f); }

__device__ __forceinline__ float loss_value(float a, float b, int loss) {
  float d = a - b;
  switch (loss) {
    case LOSS_MSE: return d * d;
    case LOSS_RELMSE: return d * d / (a * a + b * b + 0.1f);
    case LOSS_SMAPE: return fabsf(d) / (a + b + 0.01f);
    default: return fabsf(d);
  }
}
// partial derivatives of the per-channel loss w.r.t. (a, b)
__device__ __forceinline__ void loss_grad(float a, float b, int loss, float& da, float& db) {
  float d = a - b;
  switch (loss) {
    case LOSS_MSE: da = 2.f * d; db = -da; break;
    case LOSS_RELMSE: {
      float den = b * b + a * a + 0.1f, inv = 1.f / (den * den);
      da = 2.f * d * (b * (b + a) + 0.1f) * inv;
      db = -2.f * d * (a * (b + a) + 0.1f) * inv;
    } break;
    case LOSS_SMAPE: {
      float den = b + a + 0.01f, inv = 1.f / (den * den);
      da = sign0(d) * (2.f * b + 0.01f) * inv;
      db = -sign0(d) * (2.f * a + 0.01f) * inv;
    } break;
    default: da = sign0(d); db = -da;
  }
}

// partial[blockIdx.x] = sum over the block's pixels of mean_c loss(img_c, target_c)
__global__ void __launch_bounds__(kThreads) k_loss_fwd(const float* __restrict__ img, const float* __restrict__ tgt,
                                                       int64_t n_vals, int loss, int tm, float* __restrict__ partial) {
  // grid-stride over scalar channels: fully coalesced 4-byte accesses
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_vals; i += (int64_t)gridDim.x * kThreads) {
    float a = clampf(__ldg(img + i), 0.f, 65535.f), b = clampf(__ldg(tgt + i), 0.f, 65535.f);
    acc += loss_value(tonemap(a, tm), tonemap(b, tm), loss);
  }
  __shared__ float s[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += s[w];
    partial[blockIdx.x] = t / 3.f;
  }
}

__global__ void __launch_bounds__(kThreads) k_loss_bwd(const float* __restrict__ img, const float* __restrict__ tgt,
                                                       int64_t n_vals, int loss, int tm, const float* __restrict__ g_scalar,
                                                       float scale, float* __restrict__ g_img, float* __restrict__ g_tgt) {
  const float go = __ldg(g_scalar) * scale / 3.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_vals; i += (int64_t)gridDim.x * kThreads) {
    const float a0 = __ldg(img + i), b0 = __ldg(tgt + i);
    const float a = tonemap(a0, tm), b = tonemap(b0, tm);
    float da, db;
    loss_grad(a, b, loss, da, db);
    da *= go;
    db *= go;
    if (tm == TM_LOG_SRGB) {
      da *= srgb_bwd(logf(a0 + 1.f)) / (a0 + 1.f);
      db *= srgb_bwd(logf(b0 + 1.f)) / (b0 + 1.f);
    }
    // the forward clamp to [0, 65535] passes no gradient at or beyond the bounds (loss.cu:199-204)
    if (g_img) g_img[i] = (a0 <= 0.f || a0 >= 65535.f) ? 0.f : da;
    if (g_tgt) g_tgt[i] = (b0 <= 0.f || b0 >= 65535.f) ? 0.f : db;
  }
}

}  // namespace

extern "C" {

int gsb_xfm_points_fwd(const float* points, const float* matrix, int64_t n_batch, int64_t n_points,
                       int points_batched, float* out, void* stream) {
  if (n_points == 0 || n_batch == 0) return 0;
  dim3 grid(blocks_for(n_points), (unsigned)n_batch);
  k_xfm_fwd<<<grid, kThreads, 0, (cudaStream_t)stream>>>(points, matrix, n_points, points_batched, out);
  return (int)cudaGetLastError();
}

int gsb_xfm_points_bwd(const float* matrix, const float* g_out, int64_t n_batch, int64_t n_points,
                       int points_batched, float* g_points, void* stream) {
  if (n_points == 0 || n_batch == 0) return 0;
  if (n_batch * 16 * sizeof(float) > 48 * 1024) return (int)cudaErrorInvalidValue;
  dim3 grid(blocks_for(n_points), points_batched ? (unsigned)n_batch : 1u);
  k_xfm_bwd<<<grid, kThreads, n_batch * 16 * sizeof(float), (cudaStream_t)stream>>>(matrix, g_out, (int)n_batch,
                                                                                   n_points, points_batched, g_points);
  return (int)cudaGetLastError();
}

static void fill_normal_args(NormalArgs& a, const float* const* inputs, const int64_t* strides, int64_t B, int64_t H,
                             int64_t W, int two_sided, int opengl) {
  for (int k = 0; k < 6; ++k) {
    a.in[k] = inputs[k];
    a.sb[k] = strides[3 * k];
    a.sy[k] = strides[3 * k + 1];
    a.sx[k] = strides[3 * k + 2];
    a.gout[k] = nullptr;
  }
  a.B = (int)B; a.H = (int)H; a.W = (int)W; a.two_sided = two_sided; a.opengl = opengl;
  a.g = nullptr; a.out = nullptr;
}

int gsb_shading_normal_fwd(const float* const* inputs, const int64_t* strides, int64_t B, int64_t H, int64_t W,
                           int two_sided, int opengl, float* out, void* stream) {
  if (B * H * W == 0) return 0;
  NormalArgs a;
  fill_normal_args(a, inputs, strides, B, H, W, two_sided, opengl);
  a.out = out;
  k_shading_normal<false><<<blocks_for(B * H * W), kThreads, 0, (cudaStream_t)stream>>>(a);
  return (int)cudaGetLastError();
}

int gsb_shading_normal_bwd(const float* const* inputs, const int64_t* strides, int64_t B, int64_t H, int64_t W,
                           int two_sided, int opengl, const float* g_out, float* const* g_inputs, void* stream) {
  if (B * H * W == 0) return 0;
  NormalArgs a;
  fill_normal_args(a, inputs, strides, B, H, W, two_sided, opengl);
  a.g = g_out;
  for (int k = 0; k < 6; ++k) a.gout[k] = g_inputs[k];
  k_shading_normal<true><<<blocks_for(B * H * W), kThreads, 0, (cudaStream_t)stream>>>(a);
  return (int)cudaGetLastError();
}

int64_t gsb_image_loss_partials(int64_t n_values) {
  int64_t b = (n_values + kThreads * 8 - 1) / (kThreads * 8);
  if (b < 1) b = 1;
  return b > 148 * 8 ? 148 * 8 : b;
}

int gsb_image_loss_fwd(const float* img, const float* target, int64_t n_values, int loss, int tonemapper,
                       float* partials, void* stream) {
  if (loss < 0 || loss > 3 || tonemapper < 0 || tonemapper > 1) return (int)cudaErrorInvalidValue;
  int nb = (int)gsb_image_loss_partials(n_values);
  k_loss_fwd<<<nb, kThreads, 0, (cudaStream_t)stream>>>(img, target, n_values, loss, tonemapper, partials);
  return (int)cudaGetLastError();
}

int gsb_image_loss_bwd(const float* img, const float* target, int64_t n_values, int loss, int tonemapper,
                       const float* g_scalar, float scale, float* g_img, float* g_target, void* stream) {
  if (loss < 0 || loss > 3 || tonemapper < 0 || tonemapper > 1) return (int)cudaErrorInvalidValue;
  if (n_values == 0) return 0;
  int nb = (int)gsb_image_loss_partials(n_values);
  k_loss_bwd<<<nb, kThreads, 0, (cudaStream_t)stream>>>(img, target, n_values, loss, tonemapper, g_scalar, scale, g_img,
                                                       g_target);
  return (int)cudaGetLastError();
}

}  // extern "C"
