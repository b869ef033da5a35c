// Cross-bilateral denoiser (SVGF-style guides: normal^128, depth) forward and adjoint for sm_100a.
// Replaces bilateral_denoiser_fwd/bwd_kernel of the reference (render/optixutils/c_src/denoising.cu:14,74).
//
// FP32-ALU bound ((2r+1)^2 taps per pixel, r = 2*ceil(2.5 sigma)+1 = 11 at sigma 2).  B200 design:
//   * the tile (block + halo) of every guide and payload channel is staged ONCE in shared memory as
//     SoA planes (conflict-free), instead of (2r+1)^2 global fetches of 8 floats per pixel;
//   * spatial weights exp(-d^2/2 sigma^2) and tap distances come from a per-block table;
//   * pow(x,128) is seven squarings; one expf + one divide per tap remain;
//   * the renderer denoises the diffuse and the specular accumulators with identical guides
//     (render.py:140-142): both payloads go through one pass (NPAY = 2) sharing every weight.
// Output is rgb/w directly (the Python wrapper of the reference divides, ops.py:145-147) plus w, which the
// adjoint needs.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"

namespace {

constexpr int kBW = 32, kBH = 8;
constexpr float kEps = 0.0001f;

struct DenoiseArgs {
  const float* pay[2];     // fwd: colour images; bwd: d/d(out) images        [B,H,W,3] views
  const float* inv_w[2];   // bwd only: 1 / w of the forward pass              [B,H,W]
  const float* nrm;        // [B,H,W,3] view
  const float* zdz;        // [B,H,W,2] view
  int64_t ps_pay, ps_nrm, ps_zdz;   // pixel strides (floats) of the views; rows/batches are dense
  float* out[2];           // fwd: rgb / w ; bwd: d/d(col)                     [B,H,W,3] dense
  float* w_out[2];         // fwd only: max(sum w, 1e-4)                       [B,H,W]
  int B, H, W, rad;
  float inv_2var;
};

__device__ __forceinline__ float pow128(float x) {
#pragma unroll
  for (int i = 0; i < 7; ++i) x *= x;
  return x;
}

template <int NPAY, bool BWD>
__global__ void __launch_bounds__(kBW* kBH) k_bilateral(DenoiseArgs a) {
  extern __shared__ float smem[];
  const int rad = a.rad, tw = kBW + 2 * rad, th = kBH + 2 * rad, tn = tw * th, nt = 2 * rad + 1;
  float* s_n = smem;                 // 3 planes
  float* s_z = s_n + 3 * tn;         // z
  float* s_dz = s_z + tn;            // dz
  float* s_p = s_dz + tn;            // 3*NPAY planes
  float* s_wxy = s_p + 3 * NPAY * tn;   // nt*nt spatial weights
  float* s_dist = s_wxy + nt * nt;      // nt*nt distances
  const int tid = threadIdx.y * kBW + threadIdx.x;
  const int b = blockIdx.z, x0 = blockIdx.x * kBW - rad, y0 = blockIdx.y * kBH - rad;
  const int64_t img = (int64_t)b * a.H * a.W;
  for (int i = tid; i < nt * nt; i += kBW * kBH) {
    int fy = i / nt - rad, fx = i % nt - rad;
    float d2 = (float)(fx * fx + fy * fy);
    s_wxy[i] = expf(-d2 * a.inv_2var);
    s_dist[i] = sqrtf(d2);
  }
  for (int i = tid; i < tn; i += kBW * kBH) {
    int ty = i / tw, tx = i % tw, gx = x0 + tx, gy = y0 + ty;
    bool in = gx >= 0 && gx < a.W && gy >= 0 && gy < a.H;
    int64_t pix = img + (int64_t)gy * a.W + gx;
    // out-of-image taps get a zero normal => normal weight clamp(0,1e-4,1)^128 == 0 exactly
    const float* n = a.nrm + pix * a.ps_nrm;
    s_n[i] = in ? __ldg(n) : 0.f;
    s_n[tn + i] = in ? __ldg(n + 1) : 0.f;
    s_n[2 * tn + i] = in ? __ldg(n + 2) : 0.f;
    const float* z = a.zdz + pix * a.ps_zdz;
    s_z[i] = in ? __ldg(z) : 0.f;
    s_dz[i] = in ? __ldg(z + 1) : 0.f;
#pragma unroll
    for (int k = 0; k < NPAY; ++k) {
      const float* c = a.pay[k] + pix * a.ps_pay;
      float scale = 1.f;
      if (BWD && in) scale = __ldg(a.inv_w[k] + pix);     // d/d(acc) = d/d(out) / w
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) s_p[(3 * k + ch) * tn + i] = in ? __ldg(c + ch) * scale : 0.f;
    }
  }
  __syncthreads();
  const int px = blockIdx.x * kBW + threadIdx.x, py = blockIdx.y * kBH + threadIdx.y;
  if (px >= a.W || py >= a.H) return;
  const int ci = (threadIdx.y + rad) * tw + threadIdx.x + rad;
  const float cnx = s_n[ci], cny = s_n[tn + ci], cnz = s_n[2 * tn + ci], cz = s_z[ci], cdz = s_dz[ci];
  float acc[3 * NPAY];
#pragma unroll
  for (int k = 0; k < 3 * NPAY; ++k) acc[k] = 0.f;
  float acc_w = 0.f;
  for (int fy = 0; fy < nt; ++fy) {
    const int row = (threadIdx.y + fy) * tw + threadIdx.x;
    for (int fx = 0; fx < nt; ++fx) {
      const int ti = row + fx;
      const float dist = s_dist[fy * nt + fx];
      float d = s_n[ti] * cnx + s_n[tn + ti] * cny + s_n[2 * tn + ti] * cnz;
      float w_n = pow128(fminf(fmaxf(d, kEps), 1.0f));
      // forward: the CENTRE pixel's depth slope scales the tolerance (denoising.cu:59); the adjoint
      // gathers with the TAP's slope (denoising.cu:117)
      float slope = BWD ? s_dz[ti] : cdz;
      float w_z = expf(-(fabsf(s_z[ti] - cz) / fmaxf(slope * dist, kEps)));
      float w = s_wxy[fy * nt + fx] * w_n * w_z;
#pragma unroll
      for (int k = 0; k < 3 * NPAY; ++k) acc[k] += s_p[k * tn + ti] * w;
      acc_w += w;
    }
  }
  const int64_t pix = img + (int64_t)py * a.W + px;
  if (BWD) {
#pragma unroll
    for (int k = 0; k < NPAY; ++k)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) a.out[k][pix * 3 + ch] = acc[3 * k + ch];
  } else {
    const float wsum = fmaxf(acc_w, kEps), inv = 1.f / wsum;
#pragma unroll
    for (int k = 0; k < NPAY; ++k) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) a.out[k][pix * 3 + ch] = acc[3 * k + ch] * inv;
      a.w_out[k][pix] = wsum;
    }
  }
}

template <int NPAY, bool BWD>
int launch(const DenoiseArgs& a, cudaStream_t stream) {
  const int rad = a.rad, tw = kBW + 2 * rad, th = kBH + 2 * rad, nt = 2 * rad + 1;
  size_t smem = sizeof(float) * ((size_t)(5 + 3 * NPAY) * tw * th + 2 * (size_t)nt * nt);
  if (smem > 220 * 1024) return (int)cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_bilateral<NPAY, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((a.W + kBW - 1) / kBW, (a.H + kBH - 1) / kBH, a.B), block(kBW, kBH);
  k_bilateral<NPAY, BWD><<<grid, block, smem, stream>>>(a);
  return (int)cudaGetLastError();
}

int fill(DenoiseArgs& a, const float* nrm, const float* zdz, int64_t ps_pay, int64_t ps_nrm, int64_t ps_zdz, int64_t B,
         int64_t H, int64_t W, float sigma) {
  if (!(sigma > 0.f)) return (int)cudaErrorInvalidValue;
  a.nrm = nrm; a.zdz = zdz; a.ps_pay = ps_pay; a.ps_nrm = ps_nrm; a.ps_zdz = ps_zdz;
  a.B = (int)B; a.H = (int)H; a.W = (int)W;
  a.rad = 2 * (int)ceilf(sigma * 2.5f) + 1;                 // denoising.cu:28
  a.inv_2var = 1.0f / (2.0f * sigma * sigma);
  for (int k = 0; k < 2; ++k) { a.pay[k] = nullptr; a.inv_w[k] = nullptr; a.out[k] = nullptr; a.w_out[k] = nullptr; }
  return 0;
}

}  // namespace

extern "C" {

int gsb_bilateral_fwd(const float* col_a, const float* col_b, const float* nrm, const float* zdz, int64_t ps_col,
                      int64_t ps_nrm, int64_t ps_zdz, int64_t B, int64_t H, int64_t W, float sigma, float* out_a,
                      float* w_a, float* out_b, float* w_b, void* stream) {
  DenoiseArgs a;
  int err = fill(a, nrm, zdz, ps_col, ps_nrm, ps_zdz, B, H, W, sigma);
  if (err) return err;
  if (B * H * W == 0) return 0;
  a.pay[0] = col_a; a.out[0] = out_a; a.w_out[0] = w_a;
  if (!col_b) return launch<1, false>(a, (cudaStream_t)stream);
  a.pay[1] = col_b; a.out[1] = out_b; a.w_out[1] = w_b;
  return launch<2, false>(a, (cudaStream_t)stream);
}

int gsb_bilateral_bwd(const float* g_out_a, const float* w_a, const float* g_out_b, const float* w_b, const float* nrm,
                      const float* zdz, int64_t ps_nrm, int64_t ps_zdz, int64_t B, int64_t H, int64_t W, float sigma,
                      float* g_col_a, float* g_col_b, float* inv_w_scratch, void* stream);

}  // extern "C"

// 1/w planes for the adjoint (tiny elementwise kernel kept here so the library has no torch dependency)
namespace {
__global__ void k_reciprocal(const float* __restrict__ w, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = 1.f / w[i];
}
}  // namespace

extern "C" int gsb_bilateral_bwd(const float* g_out_a, const float* w_a, const float* g_out_b, const float* w_b,
                                 const float* nrm, const float* zdz, int64_t ps_nrm, int64_t ps_zdz, int64_t B,
                                 int64_t H, int64_t W, float sigma, float* g_col_a, float* g_col_b,
                                 float* inv_w_scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DenoiseArgs a;
  int err = fill(a, nrm, zdz, 3, ps_nrm, ps_zdz, B, H, W, sigma);
  if (err) return err;
  const int64_t n = B * H * W;
  if (n == 0) return 0;
  const int nb = (int)((n + 255) / 256);
  k_reciprocal<<<nb, 256, 0, stream>>>(w_a, inv_w_scratch, n);
  a.pay[0] = g_out_a; a.inv_w[0] = inv_w_scratch; a.out[0] = g_col_a;
  if (!g_out_b) return launch<1, true>(a, stream);
  k_reciprocal<<<nb, 256, 0, stream>>>(w_b, inv_w_scratch + n, n);
  a.pay[1] = g_out_b; a.inv_w[1] = inv_w_scratch + n; a.out[1] = g_col_b;
  return launch<2, true>(a, stream);
}
