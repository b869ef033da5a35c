// Loss assembly of one training iteration and the light-probe tables, as CUDA reductions with hand-written adjoints.
//
// Replaces the per-iteration PyTorch glue of the reference:
//   * EnvironmentLight.update_pdf              render/light.py:46-59            -> gsb_light_pdf
//   * compute_sdf_reg_loss                     geometry/gshell_tets_geometry.py:33-39   -> gsb_sdf_reg_fwd/bwd
//   * the two mSDF Huber regularisers          geometry/gshell_tets_geometry.py:325-356 -> gsb_msdf_reg_fwd/bwd
//   * alpha MSE, mSDF-image L1 terms           geometry/gshell_tets_geometry.py:283-290 \
//   * chroma_loss / shading_loss / material_smoothness_grad   render/regularizer.py:21-52 -> gsb_image_terms_fwd/bwd
// Every term is a mean or a sum over pixels / edges / vertices: one pass reads each buffer once and reduces into double
// accumulators (block reduction + one atomicAdd per block and term); the backward pass is elementwise and reads the
// accumulators for the terms whose gradient needs a global quantity (edge count, the specular/diffuse means).
// HBM-bound: algorithmic bytes = the buffers read once (fwd) / read once + gradients written once (bwd).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"

namespace {
constexpr int kThreads = 256;
inline int nblk(int64_t n, int cap = 148 * 8) {
  int64_t b = (n + kThreads - 1) / kThreads;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <int N>
__device__ __forceinline__ void block_reduce_add(float (&v)[N], double* __restrict__ acc) {
  __shared__ float s_part[N][kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_part[k][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += (double)s_part[threadIdx.x][w];
    if (t != 0.0) atomicAdd(acc + threadIdx.x, t);
  }
}

// ---- light probe tables -----------------------------------------------------------------------------------------------
// pass 1: one block per row: raw[y][x] = max(rgb) * sin(pi (y + 0.5) / h), row sums
__global__ void __launch_bounds__(kThreads) k_light_raw(const float* __restrict__ base, int h, int w, float* __restrict__ pdf,
                                                        float* __restrict__ row_sum) {
  const int y = blockIdx.x;
  const float sy = sinf(((float)y + 0.5f) / (float)h * 3.14159265358979323846f);
  float part[1] = {0.f};
  for (int x = threadIdx.x; x < w; x += kThreads) {
    const float* t = base + ((size_t)y * w + x) * 3;
    const float v = fmaxf(t[0], fmaxf(t[1], t[2])) * sy;
    pdf[(size_t)y * w + x] = v;
    part[0] += v;
  }
  __shared__ float s_w[kThreads / 32];
  float x = part[0];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kThreads / 32; ++k) t += s_w[k];
    row_sum[y] = t;
  }
}
// pass 2: one block per row: normalise, column CDF of the row (inclusive scan), row CDF entry
__global__ void __launch_bounds__(kThreads) k_light_tables(int h, int w, const float* __restrict__ row_sum, float* __restrict__ pdf,
                                                           float* __restrict__ cols, float* __restrict__ rows) {
  const int y = blockIdx.x;
  __shared__ float s_w[kThreads / 32];
  __shared__ float s_total, s_prefix, s_carry;
  // total of all rows and the inclusive prefix up to this row (h <= a few thousand: every block sums them itself)
  float tot = 0.f, pre = 0.f;
  for (int k = threadIdx.x; k < h; k += kThreads) {
    const float v = row_sum[k];
    tot += v;
    if (k <= y) pre += v;
  }
  for (int pass = 0; pass < 2; ++pass) {
    float x = pass == 0 ? tot : pre;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < kThreads / 32; ++k) t += s_w[k];
      if (pass == 0) s_total = t; else s_prefix = t;
    }
    __syncthreads();
  }
  const float total = s_total;
  const float inv_total = 1.f / total;
  // rows = cumsum over y of the row totals, normalised by its last entry (= sum of all normalised rows)
  const float rows_y = (s_prefix * inv_total) / ((total * inv_total) > 0.f ? (total * inv_total) : 1.f);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // two sweeps over the row: the first only finds the scan's own last value (the reference divides the column CDF by exactly
  // that, so that the last entry is 1.0), the second writes
  float norm = 1.f;
  for (int sweep = 0; sweep < 2; ++sweep) {
    if (threadIdx.x == 0) s_carry = 0.f;
    __syncthreads();
    for (int x0 = 0; x0 < w; x0 += kThreads) {
      const int x = x0 + threadIdx.x;
      const float p = x < w ? pdf[(size_t)y * w + x] * (sweep == 0 ? inv_total : 1.f) : 0.f;
      float incl = p;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      if (lane == 31) s_w[warp] = incl;
      __syncthreads();
      float off = s_carry;
      for (int k = 0; k < warp; ++k) off += s_w[k];
      if (x < w) {
        if (sweep == 0) {
          pdf[(size_t)y * w + x] = p;                                // normalised pdf (read back by the second sweep)
        } else {
          cols[(size_t)y * w + x] = (off + incl) / norm;
          rows[(size_t)y * w + x] = rows_y;
        }
      }
      __syncthreads();
      if (threadIdx.x == kThreads - 1) s_carry = off + incl;
      __syncthreads();
    }
    norm = s_carry > 0.f ? s_carry : 1.f;
    __syncthreads();
  }
}

// ---- SDF regulariser over the static edge table -------------------------------------------------------------------------
__device__ __forceinline__ float sgn(float x) { return (float)((x > 0.f) - (x < 0.f)); }
__device__ __forceinline__ float bce_logits(float x, float t) { return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kThreads) k_sdf_reg_fwd(const float* __restrict__ sdf, const int2* __restrict__ edges, int64_t E,
                                                          double* __restrict__ acc) {
  float v[2] = {0.f, 0.f};
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < E; e += (int64_t)gridDim.x * kThreads) {
    const int2 ed = __ldg(edges + e);
    const float a = __ldg(sdf + ed.x), b = __ldg(sdf + ed.y);
    if (sgn(a) != sgn(b)) {
      v[0] += bce_logits(a, b > 0.f ? 1.f : 0.f) + bce_logits(b, a > 0.f ? 1.f : 0.f);
      v[1] += 1.f;
    }
  }
  block_reduce_add<2>(v, acc);
}
__global__ void __launch_bounds__(kThreads) k_sdf_reg_bwd(const float* __restrict__ sdf, const int2* __restrict__ edges, int64_t E,
                                                          const double* __restrict__ acc, const float* __restrict__ g, float weight,
                                                          float* __restrict__ g_sdf) {
  const double cnt = acc[1];
  if (!(cnt > 0.0)) return;
  const float k = __ldg(g) * weight / (float)cnt;
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < E; e += (int64_t)gridDim.x * kThreads) {
    const int2 ed = __ldg(edges + e);
    const float a = __ldg(sdf + ed.x), b = __ldg(sdf + ed.y);
    if (sgn(a) != sgn(b)) {
      atomicAdd(g_sdf + ed.x, k * (sigmoidf(a) - (b > 0.f ? 1.f : 0.f)));
      atomicAdd(g_sdf + ed.y, k * (sigmoidf(b) - (a > 0.f ? 1.f : 0.f)));
    }
  }
}

// ---- mSDF Huber regularisers (delta = 1, reduction = sum) -------------------------------------------------------------------
__device__ __forceinline__ float huber(float d) { const float a = fabsf(d); return a < 1.f ? 0.5f * d * d : a - 0.5f; }
__device__ __forceinline__ float huber_d(float d) { return fabsf(d) < 1.f ? d : sgn(d); }

// boundary vertices (ids >= n_wt) of the listed (visible) triangles
__global__ void __launch_bounds__(kThreads) k_mark_boundary(const int32_t* __restrict__ tris, const int64_t* __restrict__ vis_ids, int64_t n_vis,
                                                            int64_t n_wt, uint8_t* __restrict__ bmask) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_vis) return;
  const int64_t f = vis_ids[i];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t v = tris[f * 3 + k];
    if (v >= n_wt) bmask[v - n_wt] = 1;
  }
}
__global__ void __launch_bounds__(kThreads) k_msdf_reg_fwd(const float* __restrict__ m_all, int64_t n_all, const float* __restrict__ m_b,
                                                           const uint8_t* __restrict__ bmask, int64_t n_b, float eps, double* __restrict__ acc) {
  float v[2] = {0.f, 0.f};
  const int64_t stride = (int64_t)gridDim.x * kThreads, t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (m_all)
    for (int64_t i = t0; i < n_all; i += stride) v[0] += huber(fmaxf(__ldg(m_all + i), -eps) + eps);     // open: pulls mSDF below -eps
  if (m_b && bmask)
    for (int64_t i = t0; i < n_b; i += stride)
      if (bmask[i]) v[1] += huber(fminf(__ldg(m_b + i), eps) - eps);                                  // close: visible boundary above eps
  block_reduce_add<2>(v, acc);
}
__global__ void __launch_bounds__(kThreads) k_msdf_reg_bwd(const float* __restrict__ m_all, int64_t n_all, const float* __restrict__ m_b,
                                                           const uint8_t* __restrict__ bmask, int64_t n_b, float eps, const float* __restrict__ g,
                                                           float w_open, float w_close, float* __restrict__ g_all, float* __restrict__ g_b) {
  const float gs = __ldg(g);
  const int64_t stride = (int64_t)gridDim.x * kThreads, t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (g_all)
    for (int64_t i = t0; i < n_all; i += stride) {
      const float m = __ldg(m_all + i);
      g_all[i] = m >= -eps ? gs * w_open * huber_d(m + eps) : 0.f;
    }
  if (g_b)
    for (int64_t i = t0; i < n_b; i += stride) {
      const float m = __ldg(m_b + i);
      g_b[i] = (bmask[i] && m <= eps) ? gs * w_close * huber_d(m - eps) : 0.f;
    }
}

// ---- image-space terms ------------------------------------------------------------------------------------------------------
// accumulators (double[kImgAcc]); T_* select terms in `terms`
enum { A_ALPHA = 0, A_MSDF_POS, A_MSDF_NEG, A_CHROMA, A_SHADE, A_SPEC, A_DIFF, A_KD, A_KS, A_NRM, kImgAcc };
enum { T_ALPHA = 1, T_MSDF = 2, T_CHROMA = 4, T_SHADING = 8, T_SMOOTH = 16 };

struct ImgArgs {
  const float *shaded, *msdf_img, *kd, *kd_grad, *ks_grad, *nrm_grad, *diffuse, *specular, *ref;   // [N,4] (msdf_img [N,msdf_ch]); null = absent
  int64_t n_pix;
  int msdf_ch, terms;
  float lambda_chroma, lambda_diffuse, lambda_specular, lambda_kd, lambda_ks, lambda_nrm;
};
struct ImgGrads {
  float *shaded, *msdf_img, *kd, *kd_grad, *ks_grad, *nrm_grad, *diffuse, *specular;               // same shapes; null = not wanted
};

__device__ __forceinline__ float srgb(float f) { return f <= 0.0031308f ? f * 12.92f : powf(fmaxf(f, 0.0031308f), 1.0f / 2.4f) * 1.055f - 0.055f; }
__device__ __forceinline__ float srgb_d(float f) { return f <= 0.0031308f ? 12.92f : (1.055f / 2.4f) * powf(f, 1.0f / 2.4f - 1.f); }
__device__ __forceinline__ float4 ld4(const float* p, int64_t i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }

__global__ void __launch_bounds__(kThreads) k_img_terms_fwd(ImgArgs a, double* __restrict__ acc) {
  float v[kImgAcc];
#pragma unroll
  for (int k = 0; k < kImgAcc; ++k) v[k] = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < a.n_pix; i += (int64_t)gridDim.x * kThreads) {
    const float4 r = ld4(a.ref, i);
    if ((a.terms & T_ALPHA) && a.shaded) {
      const float d = __ldg(a.shaded + i * 4 + 3) - r.w;
      v[A_ALPHA] += d * d;
    }
    if ((a.terms & T_MSDF) && a.msdf_img) {
      for (int c = 0; c < a.msdf_ch; ++c) {
        const float x = __ldg(a.msdf_img + i * a.msdf_ch + c);
        v[A_MSDF_POS] += fabsf(fmaxf(x, 0.f) * (r.w == 0.f ? 1.f : 0.f));
        v[A_MSDF_NEG] += fabsf(fminf(x, 0.f) * (r.w == 1.f ? 1.f : 0.f) - 1.f);
      }
    }
    if ((a.terms & T_CHROMA) && a.kd) {
      const float4 k = ld4(a.kd, i);
      const float vk = fmaxf(fmaxf(k.x, fmaxf(k.y, k.z)), 0.001f), vr = fmaxf(fmaxf(r.x, fmaxf(r.y, r.z)), 0.001f);
      v[A_CHROMA] += fabsf((k.x / vk - r.x / vr) * r.w) + fabsf((k.y / vk - r.y / vr) * r.w) + fabsf((k.z / vk - r.z / vr) * r.w);
    }
    if ((a.terms & T_SHADING) && a.diffuse && a.specular) {
      const float4 d = ld4(a.diffuse, i), s = ld4(a.specular, i);
      const float dl = (d.x + d.y + d.z) / 3.f, sl = (s.x + s.y + s.z) / 3.f, vr = fmaxf(r.x, fmaxf(r.y, r.z));
      const float img = srgb(logf(fminf(fmaxf((dl + sl) * r.w, 0.f), 65535.f) + 1.f));
      const float tgt = srgb(logf(fminf(fmaxf(vr * r.w, 0.f), 65535.f) + 1.f));
      v[A_SHADE] += fabsf(img - tgt);
      v[A_SPEC] += sl;
      v[A_DIFF] += dl;
    }
    if (a.terms & T_SMOOTH) {
      if (a.kd_grad) { const float4 k = ld4(a.kd_grad, i); v[A_KD] += (k.x + k.y + k.z) / 3.f * k.w; }
      if (a.ks_grad) { const float4 k = ld4(a.ks_grad, i); v[A_KS] += (k.x + k.y + k.z) * k.w; }
      if (a.nrm_grad) { const float4 k = ld4(a.nrm_grad, i); v[A_NRM] += (k.x + k.y + k.z) * k.w; }
    }
  }
  block_reduce_add<kImgAcc>(v, acc);
}

// out[0] = image part (alpha MSE + 0.5 * both mSDF-image L1 terms), out[1] = regulariser part (chroma + shading + smoothness).
// mean_scale: 1 / world size when acc[A_SPEC], acc[A_DIFF] were summed over ranks of equal pixel counts (else 1).
__global__ void k_img_terms_finish(ImgArgs a, const double* __restrict__ acc, float mean_scale, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double n = (double)a.n_pix;
  double img = 0.0, reg = 0.0;
  if (a.terms & T_ALPHA) img += acc[A_ALPHA] / n;
  if ((a.terms & T_MSDF) && a.msdf_img) img += 0.5 * (acc[A_MSDF_POS] + acc[A_MSDF_NEG]) / (n * a.msdf_ch);
  if (a.terms & T_CHROMA) reg += acc[A_CHROMA] / (3.0 * n) * a.lambda_chroma;
  if ((a.terms & T_SHADING) && a.diffuse && a.specular) {
    reg += acc[A_SHADE] / n * a.lambda_diffuse;
    const double ms = acc[A_SPEC] * mean_scale / n, md = acc[A_DIFF] * mean_scale / n;
    reg += ms / (md > 0.001 ? md : 0.001) * a.lambda_specular;
  }
  if (a.terms & T_SMOOTH) reg += acc[A_KD] / n * a.lambda_kd + acc[A_KS] / (3.0 * n) * a.lambda_ks + acc[A_NRM] / (3.0 * n) * a.lambda_nrm;
  out[0] = (float)img;
  out[1] = (float)reg;
}

__device__ __forceinline__ void st4(float* p, int64_t i, float x, float y, float z, float w) {
  reinterpret_cast<float4*>(p)[i] = make_float4(x, y, z, w);
}

__global__ void __launch_bounds__(kThreads) k_img_terms_bwd(ImgArgs a, const double* __restrict__ acc, float mean_scale, const float* __restrict__ g_out,
                                                            ImgGrads go) {
  const float gi = __ldg(g_out), gr = __ldg(g_out + 1);
  const float n = (float)a.n_pix, inv_n = 1.f / n;
  const float ms = (float)(acc[A_SPEC] * mean_scale) * inv_n, md = (float)(acc[A_DIFF] * mean_scale) * inv_n;
  const float mdc = fmaxf(md, 0.001f);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < a.n_pix; i += (int64_t)gridDim.x * kThreads) {
    const float4 r = ld4(a.ref, i);
    if (go.shaded) {
      const float d = (a.terms & T_ALPHA) ? 2.f * (__ldg(a.shaded + i * 4 + 3) - r.w) * inv_n * gi : 0.f;
      st4(go.shaded, i, 0.f, 0.f, 0.f, d);
    }
    if (go.msdf_img) {
      const float k = (a.terms & T_MSDF) ? 0.5f * gi / (n * a.msdf_ch) : 0.f;
      for (int c = 0; c < a.msdf_ch; ++c) {
        const float x = __ldg(a.msdf_img + i * a.msdf_ch + c);
        // |max(x,0) m0| has slope m0 for x > 0;  |min(x,0) m1 - 1| = 1 - min(x,0) m1 has slope -m1 for x < 0
        go.msdf_img[i * a.msdf_ch + c] = k * ((x > 0.f && r.w == 0.f ? 1.f : 0.f) - (x < 0.f && r.w == 1.f ? 1.f : 0.f));
      }
    }
    if (go.kd) {
      float gx = 0.f, gy = 0.f, gz = 0.f;
      if (a.terms & T_CHROMA) {
        const float4 k = ld4(a.kd, i);
        const float vmax = fmaxf(k.x, fmaxf(k.y, k.z)), vk = fmaxf(vmax, 0.001f), vr = fmaxf(fmaxf(r.x, fmaxf(r.y, r.z)), 0.001f);
        const float c = a.lambda_chroma * gr * inv_n / 3.f * r.w;
        const float sx = c * sgn((k.x / vk - r.x / vr) * r.w), sy = c * sgn((k.y / vk - r.y / vr) * r.w), sz = c * sgn((k.z / vk - r.z / vr) * r.w);
        gx = sx / vk; gy = sy / vk; gz = sz / vk;
        if (vmax >= 0.001f) {                                       // through the max: to the (first) largest channel
          const float gv = -(sx * k.x + sy * k.y + sz * k.z) / (vk * vk);
          if (k.x >= k.y && k.x >= k.z) gx += gv; else if (k.y >= k.z) gy += gv; else gz += gv;
        }
      }
      st4(go.kd, i, gx, gy, gz, 0.f);
    }
    if (go.diffuse || go.specular) {
      float gd = 0.f, gsp = 0.f;
      if (a.terms & T_SHADING) {
        const float4 d = ld4(a.diffuse, i), s = ld4(a.specular, i);
        const float dl = (d.x + d.y + d.z) / 3.f, sl = (s.x + s.y + s.z) / 3.f, vr = fmaxf(r.x, fmaxf(r.y, r.z));
        const float raw = (dl + sl) * r.w;
        const float u = fminf(fmaxf(raw, 0.f), 65535.f), w = logf(u + 1.f);
        const float tgt = srgb(logf(fminf(fmaxf(vr * r.w, 0.f), 65535.f) + 1.f));
        float t = 0.f;
        if (raw >= 0.f && raw <= 65535.f) t = sgn(srgb(w) - tgt) * srgb_d(w) / (u + 1.f) * r.w * a.lambda_diffuse * inv_n;
        gd = t + (md >= 0.001f ? -ms / (mdc * mdc) * a.lambda_specular * inv_n : 0.f);
        gsp = t + a.lambda_specular * inv_n / mdc;
        gd *= gr / 3.f;
        gsp *= gr / 3.f;
      }
      if (go.diffuse) st4(go.diffuse, i, gd, gd, gd, 0.f);
      if (go.specular) st4(go.specular, i, gsp, gsp, gsp, 0.f);
    }
    if (go.kd_grad) {
      const float k = (a.terms & T_SMOOTH) ? __ldg(a.kd_grad + i * 4 + 3) * a.lambda_kd * inv_n / 3.f * gr : 0.f;
      st4(go.kd_grad, i, k, k, k, 0.f);
    }
    if (go.ks_grad) {
      const float k = (a.terms & T_SMOOTH) ? __ldg(a.ks_grad + i * 4 + 3) * a.lambda_ks * inv_n / 3.f * gr : 0.f;
      st4(go.ks_grad, i, k, k, k, 0.f);
    }
    if (go.nrm_grad) {
      const float k = (a.terms & T_SMOOTH) ? __ldg(a.nrm_grad + i * 4 + 3) * a.lambda_nrm * inv_n / 3.f * gr : 0.f;
      st4(go.nrm_grad, i, k, k, k, 0.f);
    }
  }
}

ImgArgs make_args(const float* shaded, const float* msdf_img, const float* kd, const float* kd_grad, const float* ks_grad,
                  const float* nrm_grad, const float* diffuse, const float* specular, const float* ref, int64_t n_pix, int msdf_ch,
                  int terms, const float* lambdas) {
  ImgArgs a;
  a.shaded = shaded; a.msdf_img = msdf_img; a.kd = kd; a.kd_grad = kd_grad; a.ks_grad = ks_grad; a.nrm_grad = nrm_grad;
  a.diffuse = diffuse; a.specular = specular; a.ref = ref; a.n_pix = n_pix; a.msdf_ch = msdf_ch; a.terms = terms;
  a.lambda_chroma = lambdas[0]; a.lambda_diffuse = lambdas[1]; a.lambda_specular = lambdas[2];
  a.lambda_kd = lambdas[3]; a.lambda_ks = lambdas[4]; a.lambda_nrm = lambdas[5];
  return a;
}

}  // namespace

extern "C" {

int gsb_light_pdf(const float* base, int64_t h, int64_t w, float* row_sum_ws, float* pdf, float* cols, float* rows, void* stream_) {
  if (h < 1 || w < 1 || h > 65535) return (int)cudaErrorInvalidValue;
  cudaStream_t stream = (cudaStream_t)stream_;
  k_light_raw<<<(unsigned)h, kThreads, 0, stream>>>(base, (int)h, (int)w, pdf, row_sum_ws);
  k_light_tables<<<(unsigned)h, kThreads, 0, stream>>>((int)h, (int)w, row_sum_ws, pdf, cols, rows);
  return (int)cudaGetLastError();
}

int gsb_sdf_reg_fwd(const float* sdf, const int32_t* edges, int64_t n_edges, double* acc2, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(acc2, 0, 2 * sizeof(double), stream);
  if (e != cudaSuccess) return (int)e;
  if (n_edges > 0) k_sdf_reg_fwd<<<nblk(n_edges), kThreads, 0, stream>>>(sdf, (const int2*)edges, n_edges, acc2);
  return (int)cudaGetLastError();
}
int gsb_sdf_reg_bwd(const float* sdf, const int32_t* edges, int64_t n_edges, const double* acc2, const float* g_loss, float weight,
                    float* g_sdf, void* stream_) {
  if (n_edges > 0)
    k_sdf_reg_bwd<<<nblk(n_edges), kThreads, 0, (cudaStream_t)stream_>>>(sdf, (const int2*)edges, n_edges, acc2, g_loss, weight, g_sdf);
  return (int)cudaGetLastError();
}

int gsb_mark_visible_boundary(const int32_t* tris, const int64_t* visible_ids, int64_t n_visible, int64_t n_verts_watertight,
                              uint8_t* bmask, void* stream_) {
  if (n_visible > 0)
    k_mark_boundary<<<(unsigned)((n_visible + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream_>>>(tris, visible_ids, n_visible,
                                                                                                            n_verts_watertight, bmask);
  return (int)cudaGetLastError();
}
int gsb_msdf_reg_fwd(const float* msdf_all, int64_t n_all, const float* msdf_boundary, const uint8_t* bmask, int64_t n_boundary, float eps,
                     double* acc2, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(acc2, 0, 2 * sizeof(double), stream);
  if (e != cudaSuccess) return (int)e;
  const int64_t n = n_all > n_boundary ? n_all : n_boundary;
  if (n > 0) k_msdf_reg_fwd<<<nblk(n), kThreads, 0, stream>>>(msdf_all, n_all, msdf_boundary, bmask, n_boundary, eps, acc2);
  return (int)cudaGetLastError();
}
int gsb_msdf_reg_bwd(const float* msdf_all, int64_t n_all, const float* msdf_boundary, const uint8_t* bmask, int64_t n_boundary, float eps,
                     const float* g_loss, float w_open, float w_close, float* g_all, float* g_boundary, void* stream_) {
  const int64_t n = n_all > n_boundary ? n_all : n_boundary;
  if (n > 0)
    k_msdf_reg_bwd<<<nblk(n), kThreads, 0, (cudaStream_t)stream_>>>(msdf_all, n_all, msdf_boundary, bmask, n_boundary, eps, g_loss, w_open,
                                                                   w_close, g_all, g_boundary);
  return (int)cudaGetLastError();
}

int gsb_image_terms_accumulators(void) { return kImgAcc; }

int gsb_image_terms_reduce(const float* shaded, const float* msdf_img, const float* kd, const float* kd_grad, const float* ks_grad,
                           const float* nrm_grad, const float* diffuse, const float* specular, const float* ref, int64_t n_pix, int msdf_ch,
                           int terms, const float* lambdas6, double* acc, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(acc, 0, kImgAcc * sizeof(double), stream);
  if (e != cudaSuccess) return (int)e;
  if (n_pix > 0)
    k_img_terms_fwd<<<nblk(n_pix), kThreads, 0, stream>>>(make_args(shaded, msdf_img, kd, kd_grad, ks_grad, nrm_grad, diffuse, specular, ref,
                                                                    n_pix, msdf_ch, terms, lambdas6), acc);
  return (int)cudaGetLastError();
}
int gsb_image_terms_finish(const float* msdf_img, const float* diffuse, const float* specular, int64_t n_pix, int msdf_ch, int terms,
                           const float* lambdas6, const double* acc, float mean_scale, float* out2, void* stream_) {
  k_img_terms_finish<<<1, 32, 0, (cudaStream_t)stream_>>>(make_args(nullptr, msdf_img, nullptr, nullptr, nullptr, nullptr, diffuse, specular,
                                                                    nullptr, n_pix, msdf_ch, terms, lambdas6), acc, mean_scale, out2);
  return (int)cudaGetLastError();
}
int gsb_image_terms_bwd(const float* shaded, const float* msdf_img, const float* kd, const float* kd_grad, const float* ks_grad,
                        const float* nrm_grad, const float* diffuse, const float* specular, const float* ref, int64_t n_pix, int msdf_ch,
                        int terms, const float* lambdas6, const double* acc, float mean_scale, const float* g_out2, float* g_shaded,
                        float* g_msdf_img, float* g_kd, float* g_kd_grad, float* g_ks_grad, float* g_nrm_grad, float* g_diffuse,
                        float* g_specular, void* stream_) {
  ImgGrads go;
  go.shaded = g_shaded; go.msdf_img = g_msdf_img; go.kd = g_kd; go.kd_grad = g_kd_grad; go.ks_grad = g_ks_grad; go.nrm_grad = g_nrm_grad;
  go.diffuse = g_diffuse; go.specular = g_specular;
  if (n_pix > 0)
    k_img_terms_bwd<<<nblk(n_pix), kThreads, 0, (cudaStream_t)stream_>>>(make_args(shaded, msdf_img, kd, kd_grad, ks_grad, nrm_grad, diffuse,
                                                                                   specular, ref, n_pix, msdf_ch, terms, lambdas6),
                                                                         acc, mean_scale, g_out2, go);
  return (int)cudaGetLastError();
}

}  // extern "C"
