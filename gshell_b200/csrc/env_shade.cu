// Monte-Carlo environment-light integrator with MIS (light-importance + BSDF-importance samples),
// forward and sample-replaying analytic backward, for sm_100a.
//
// Replaces the OptiX raygen program of the reference (render/optixutils/c_src/envsampling/kernel.cu:
// 463-542 `__raygen__rg`, process_sample :403-461, light/BSDF sampling :124-397, BSDF fwd/bwd
// render/optixutils/c_src/bsdf.h:21-276) and its host launch (torch_bindings.cpp:123-272) with a plain
// CUDA kernel: no OptiX, no NVRTC, no per-call cudaMalloc / stream sync.
//
// FP32-ALU/SFU-bound (2 n^2 BSDF evaluations + 2 n^2 CDF searches per covered pixel): one thread per
// pixel, G-buffer rows read once, per-pixel gradients accumulated in registers and written once
// (the reference does a global read-modify-write per sample, kernel.cu:442-456), light gradient
// scattered with red.global.add.  Sample set, PCG streams, strata permutation, pdfs and MIS weights
// follow the reference exactly so that results are comparable sample by sample.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

#include "vec.cuh"

using namespace gsb;

namespace {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kEps = 1e-4f;          // SPECULAR_EPSILON (bsdf.h:13)
constexpr float kMinRough = 0.08f;     // MIN_ROUGHNESS (kernel.cu:17)

struct ShadeParams {
  const float *mask, *ro, *pos, *nrm, *view_pos, *kd, *ks;       // G-buffer ([B,H,W](,3); view_pos [B,3])
  const float *light, *pdf, *rows, *cols;                        // probe [lh,lw,3], pdf [lh,lw], cdfs
  const float *rows_top, *cols_top;                              // [16], [lh,16]: every 16th CDF entry (2.0 padded) or null
  const int32_t* perms;                                          // [n_perms, n*n]
  const float *g_diff, *g_spec;                                  // backward inputs
  float *diff, *spec;                                            // forward outputs
  float *g_pos, *g_nrm, *g_kd, *g_ks, *g_light;                  // backward outputs
  float4* ray_list;                                              // GEN: compact list of shadow rays, 2 float4 each: (origin, ray id), (direction, 0)
  int* ray_count;                                                // GEN: device counter of list entries
  int ray_cap;                                                   // GEN: capacity of the list (an entry past it is counted in *dropped)
  const uint32_t* pixel_ids;                                     // optional [B*H*W]: the number that seeds a pixel's sample stream instead of its index
  unsigned int* dropped;                                         // GEN: rays that did not fit (caller's n_covered was not an upper bound)
  const uint8_t* vis_chunk;                                      // FWD/BWD: [2 (i1-i0)][B*H*W] visibility of this chunk's rays, or null
  uint32_t* vis_out;                                             // FWD: optional [B*H*W, vis_words] visibility bits of every sample
  const uint32_t* vis_in;                                        // BWD: optional, replays the forward's bits instead of vis_chunk
  int vis_words;
  int i0, i1, first_chunk;                                       // sample-pair range of this launch; first_chunk => overwrite outputs
  int B, H, W, lh, lw, n_perms, bsdf, n;
  uint32_t seed;
  float shadow_scale;
};

// ---- PCG (kernel.cu:30-45) -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t pcg_next(uint32_t& s) {
  uint32_t word = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
  s = s * 747796405u + 2891336453u;
  return (word >> 22u) ^ word;
}
__device__ __forceinline__ uint32_t pcg_hash(uint32_t a, uint32_t b) { return pcg_next(a) ^ pcg_next(b); }
__device__ __forceinline__ float pcg_uniform(uint32_t& s) { return (float)(pcg_next(s) & 0xFFFFFFu) / (float)0x1000000; }

// ---- lat-long probe (kernel.cu:124-211) ----------------------------------------------------------
__device__ __forceinline__ void dir_to_tc(V3 d, float& u, float& v) {
  u = atan2f(d.x, -d.z) / (2.0f * kPi) + 0.5f;
  v = acosf(clampf(d.y, -1.f, 1.f)) / kPi;
}
__device__ __forceinline__ int texel_index(const ShadeParams& p, float u, float v) {
  int x = min(max((int)(u * p.lw), 0), p.lw - 1);
  int y = min(max((int)(v * p.lh), 0), p.lh - 1);
  return y * p.lw + x;
}
// binary search of a normalised CDF (kernel.cu:140-169); returns the remapped sample in [0,1)
__device__ __forceinline__ float sample_cdf(const float* __restrict__ cdf, int n, int n_steps, float x, int& idx) {
  x = fminf(x, 0.99999994f);
  int lo = 0, hi = n - 1;
  for (int i = 0; i < n_steps; ++i) {
    int mid = (lo + hi) >> 1;
    float c = __ldg(cdf + mid);
    lo = x >= c ? mid : lo;
    hi = x < c ? mid : hi;
  }
  idx = hi;
  float pdf, s;
  if (hi == 0) {
    pdf = __ldg(cdf);
    s = x;
  } else {
    float d1 = __ldg(cdf + hi - 1);
    pdf = __ldg(cdf + hi) - d1;
    s = x - d1;
  }
  return fminf(s / pdf, 0.99999994f);
}
// Same result as sample_cdf (idx = first entry with x < cdf[idx], else n-1) with two dependent 64-byte loads instead
// of ceil(log2 n)+1 dependent 4-byte loads: `top` holds cdf[15], cdf[31], ... (padded with 2.0), n % 16 == 0, n <= 256.
__device__ __forceinline__ int count_ge(const float4* __restrict__ q, float x) {
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 v = __ldg(q + k);
    c += (x >= v.x) + (x >= v.y) + (x >= v.z) + (x >= v.w);
  }
  return c;
}
__device__ __forceinline__ float sample_cdf16(const float* __restrict__ cdf, const float* __restrict__ top, int n, float x,
                                              int& idx) {
  x = fminf(x, 0.99999994f);
  const int nb = n >> 4;
  const int bucket = min(count_ge(reinterpret_cast<const float4*>(top), x), nb - 1);
  const int j = count_ge(reinterpret_cast<const float4*>(cdf + 16 * bucket), x);
  const int hi = min(16 * bucket + j, n - 1);
  idx = hi;
  float pdf, sm;                          // the two entries below were just brought in: L1 hits
  if (hi == 0) {
    pdf = __ldg(cdf);
    sm = x;
  } else {
    const float lo = __ldg(cdf + hi - 1);
    pdf = __ldg(cdf + hi) - lo;
    sm = x - lo;
  }
  return fminf(sm / pdf, 0.99999994f);
}
// lightPDF (kernel.cu:171-182) for an arbitrary direction; also returns the probe texel of `d`
__device__ __forceinline__ float light_pdf(const ShadeParams& p, V3 d, int& tex) {
  float u, v;
  dir_to_tc(d, u, v);
  tex = texel_index(p, u, v);
  float w = (float)(p.lh * p.lw) / (2.0f * kPi * kPi * fmaxf(sinf(v * kPi), 0.0001f));
  return __ldg(p.pdf + tex) * w;
}
// lightSample (kernel.cu:184-193).  The reference maps the sampled (u,v) to a direction and back to (u,v) to look up
// pdf and radiance; the round trip lands in the sampled texel (x,y) and sin(v pi) is the sin(theta) already computed,
// so both lookups use (x,y) directly (identical up to one-ulp ties on a texel border).
__device__ __forceinline__ V3 light_sample(const ShadeParams& p, int steps_r, int steps_c, float su, float sv, float& pdf,
                                           int& tex) {
  int x, y;
  float ry, rx;
  if (p.rows_top) {
    ry = sample_cdf16(p.rows, p.rows_top, p.lh, sv, y);
    rx = sample_cdf16(p.cols + (size_t)y * p.lw, p.cols_top + (size_t)y * 16, p.lw, su, x);
  } else {
    ry = sample_cdf(p.rows, p.lh, steps_r, sv, y);
    rx = sample_cdf(p.cols + (size_t)y * p.lw, p.lw, steps_c, su, x);
  }
  const float u = ((float)x + rx) / (float)p.lw, v = ((float)y + ry) / (float)p.lh;
  float sp, cp, st, ct;
  sincosf((u * 2.f - 1.f) * kPi, &sp, &cp);
  sincosf(v * kPi, &st, &ct);
  tex = y * p.lw + x;
  pdf = __ldg(p.pdf + tex) * ((float)(p.lh * p.lw) / (2.0f * kPi * kPi * fmaxf(st, 0.0001f)));
  return v3(st * sp, ct, -st * cp);
}

// ---- BSDF importance sampling (kernel.cu:217-397) ------------------------------------------------
__device__ __forceinline__ float ndf_s(float alpha, float c) {   // evalNdfGGX (:217-222), unclamped
  float a2 = alpha * alpha;
  float d = (c * a2 - c) * c + 1.f;
  return a2 / (d * d * kPi);
}
__device__ __forceinline__ float g1_s(float a2, float c) {       // evalG1GGX (:224-230)
  if (c <= 0.f) return 0.f;
  float c2 = c * c;
  float t2 = fmaxf(1.f - c2, 0.f) / c2;
  return 2.f / (1.f + sqrtf(1.f + a2 * t2));
}
struct Frame {
  V3 u, v, w;
};
__device__ __forceinline__ Frame make_frame(V3 n) {
  Frame f;
  f.w = normalize0(n);
  onb(f.w, f.u, f.v);
  return f;
}
__device__ __forceinline__ V3 to_local(const Frame& f, V3 a) { return v3(dot(a, f.u), dot(a, f.v), dot(a, f.w)); }
__device__ __forceinline__ V3 to_world(const Frame& f, V3 a) { return f.u * a.x + f.v * a.y + f.w * a.z; }

__device__ __forceinline__ float ggx_pdf(const Frame& f, V3 wo, V3 wi, float alpha) {   // :301-323
  // only rotation-invariant quantities of the local frame are needed: z components and the half vector
  const float zo = dot(wo, f.w), zi = dot(wi, f.w);
  if (!(zo > 0.f && zi > 0.f)) return 0.f;
  const V3 h = normalize0(wi + wo);
  const float wo_h = dot(h, wo);
  float pdf = g1_s(alpha * alpha, zo) * ndf_s(alpha, dot(h, f.w)) * fmaxf(0.f, wo_h) / zo;
  return pdf / (4.f * wo_h);
}
__device__ __forceinline__ void mix_pdf(float& pdf, float other, float b) {             // update_pdf :325-332
  if (b > 0.000001f) pdf += other * b;
}
__device__ __forceinline__ float bsdf_pdf(const Frame& f, float p_d, float p_s, V3 n, V3 wo, V3 wi, float alpha) {
  float n_l = dot(n, wi), n_v = dot(n, wo);
  if (fminf(n_v, n_l) < 1e-6f) return 1.0f;                                             // :382-383
  float pdf = 0.f;
  if (p_d > 0.f) mix_pdf(pdf, fmaxf(n_l, 0.f) / kPi, p_d);
  if (p_s > 0.f) mix_pdf(pdf, ggx_pdf(f, wo, wi, alpha), 1.f - p_d);
  return pdf;
}
__device__ __forceinline__ V3 cosine_sample(const Frame& f, float u, float v, float& pdf) {   // :57-79
  float sp, cp;
  sincosf(2.0f * kPi * u, &sp, &cp);
  float ct = sqrtf(v), st = sqrtf(1.0f - v);
  pdf = fmaxf(0.000001f, ct / kPi);
  return normalize0(f.u * (cp * st) + f.v * (sp * st) + f.w * ct);
}
__device__ __forceinline__ V3 ggx_sample(const Frame& f, V3 wo, float ux, float uy, float alpha, float& pdf) {  // :241-291
  V3 wo_l = normalize0(to_local(f, wo));
  if (!(wo_l.z > 0.f)) {
    pdf = 0.f;
    return v3(0.f);
  }
  V3 vh = normalize0(v3(alpha * wo_l.x, alpha * wo_l.y, wo_l.z));
  V3 t1 = vh.z < 0.9999f ? normalize0(cross(v3(0.f, 0.f, 1.f), vh)) : v3(1.f, 0.f, 0.f);
  V3 t2 = cross(vh, t1);
  float r = sqrtf(ux), sp, cp;
  sincosf(2.f * kPi * uy, &sp, &cp);
  float a = r * cp, b = r * sp;
  float s = 0.5f * (1.f + vh.z);
  b = (1.f - s) * sqrtf(1.f - a * a) + s * b;
  V3 nh = t1 * a + t2 * b + vh * sqrtf(fmaxf(0.f, 1.f - a * a - b * b));
  V3 h = normalize0(v3(alpha * nh.x, alpha * nh.y, fmaxf(0.f, nh.z)));
  pdf = g1_s(alpha * alpha, wo_l.z) * ndf_s(alpha, h.z) * fmaxf(0.f, dot(wo_l, h)) / wo_l.z;
  float wo_h = dot(wo_l, h);
  V3 wi_l = h * (wo_h * 2.f) - wo_l;
  pdf /= 4.f * wo_h;
  return normalize0(to_world(f, wi_l));
}
__device__ __forceinline__ V3 bsdf_sample(const Frame& f, float p_d, float p_s, V3 n, V3 wo, float sx, float sy, float sz,
                                          float alpha, float& pdf) {                     // :334-372
  V3 wi;
  if (sz < p_d) {
    if (p_d < 0.0001f) {
      pdf = 1.f;
      return n;
    }
    wi = cosine_sample(f, sx, sy, pdf);
    pdf *= p_d;
    if (p_s > 0.f) mix_pdf(pdf, ggx_pdf(f, wo, wi, alpha), 1.f - p_d);
  } else {
    wi = ggx_sample(f, wo, sx, sy, alpha, pdf);
    pdf *= 1.f - p_d;
    if (p_d > 0.f) mix_pdf(pdf, fmaxf(dot(n, wi), 0.f) / kPi, p_d);
  }
  return wi;
}

// ---- BSDF evaluation (bsdf.h) ---------------------------------------------------------------------
__device__ __forceinline__ float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }
__device__ __forceinline__ float lambda_ggx(float a2, float cos_t) {
  float c = clampf(cos_t, kEps, 1.f - kEps), c2 = c * c;
  return 0.5f * (sqrtf(1.f + a2 * (1.f - c2) / c2) - 1.f);
}

struct SurfaceConst {      // per-pixel constants of the specular lobe
  V3 n, wo, wo_raw, spec_col, kd, arm;
  float alpha;             // arm.y^2 (unclamped)
};

// demodulated diffuse (Lambert, no kd) + GGX specular (bsdf.h:222-236, :144-162)
__device__ __forceinline__ void eval_bsdf(const SurfaceConst& s, V3 wi, int bsdf, float& diff, V3& spec) {
  diff = fmaxf(dot(s.n, wi) / kPi, 0.f);
  spec = v3(0.f);
  if (bsdf != 0) return;
  float wo_n = dot(s.wo, s.n), wi_n = dot(wi, s.n);
  if (!(wo_n > kEps && wi_n > kEps)) return;
  float a = clampf(s.alpha, kMinRough * kMinRough, 1.f), a2 = a * a;
  V3 h = normalize0(s.wo + wi);
  float wo_h = dot(s.wo, h), n_h = dot(s.n, h);
  float c = clampf(n_h, kEps, 1.f - kEps);
  float dd = (c * a2 - c) * c + 1.f;
  float D = a2 / (dd * dd * kPi);
  float G = 1.f / (1.f + lambda_ggx(a2, wo_n) + lambda_ggx(a2, wi_n));
  float sc = pow5(1.f - clampf(wo_h, kEps, 1.f - kEps));
  V3 F = s.spec_col * (1.f - sc) + v3(sc);
  spec = F * (D * G * 0.25f / wo_n);
}

struct PixelGrads {
  V3 kd, arm, pos, nrm;
};

// adjoint of eval_bsdf w.r.t. (kd, arm, pos, nrm); g_diff is d/d(diff) summed over channels (bsdf.h:164-275)
__device__ __forceinline__ void eval_bsdf_bwd(const SurfaceConst& s, V3 wi, int bsdf, float g_diff, V3 g_spec,
                                              PixelGrads& out) {
  if (dot(s.n, wi) > 0.f) out.nrm += wi * (g_diff / kPi);
  if (bsdf != 0) return;
  float wo_n = dot(s.wo, s.n), wi_n = dot(wi, s.n);
  if (!(wo_n > kEps && wi_n > kEps)) return;
  float a = clampf(s.alpha, kMinRough * kMinRough, 1.f), a2 = a * a;
  V3 h_raw = s.wo + wi, h = normalize0(h_raw);
  float wo_h = dot(s.wo, h), n_h = dot(s.n, h);
  float c = clampf(n_h, kEps, 1.f - kEps), c2 = c * c;
  float dd = (c * a2 - c) * c + 1.f;
  float D = a2 / (dd * dd * kPi);
  float lam_o = lambda_ggx(a2, wo_n), lam_i = lambda_ggx(a2, wi_n);
  float G = 1.f / (1.f + lam_o + lam_i);
  float sc = pow5(1.f - clampf(wo_h, kEps, 1.f - kEps));
  V3 F = s.spec_col * (1.f - sc) + v3(sc);
  float k = 0.25f / wo_n;
  V3 gF = g_spec * (D * G * k);
  float gD = dot(g_spec, F) * (G * k);
  float gG = dot(g_spec, F) * (D * k);
  float g_wo_n = -dot(g_spec, F) * (D * G * k / wo_n);
  float g_wo_h = 0.f, g_wi_n = 0.f, g_n_h = 0.f, g_a2 = 0.f;
  // Fresnel-Schlick with f90 = 1
  V3 g_col = gF * (1.f - sc);
  if (wo_h >= kEps && wo_h < 1.f - kEps) {
    float q = 1.f - wo_h, q2 = q * q;
    g_wo_h += dot(gF, v3(1.f) - s.spec_col) * (-5.f * q2 * q2);
  }
  // Smith masking
  {
    float gl = -gG * G * G;
    auto lam_bwd = [&](float cos_t, float& g_cos) {
      float cc = clampf(cos_t, kEps, 1.f - kEps), cc2 = cc * cc;
      float t2 = (1.f - cc2) / cc2;
      g_a2 += gl * (0.25f * t2) / sqrtf(a2 * t2 + 1.f);
      if (cos_t > kEps && cos_t < 1.f - kEps) g_cos += gl * -(0.5f * a2) / (cc * cc2 * sqrtf(a2 / cc2 - a2 + 1.f));
    };
    lam_bwd(wo_n, g_wo_n);
    lam_bwd(wi_n, g_wi_n);
  }
  // GGX normal distribution
  {
    float base = (a2 - 1.f) * c2 + 1.f, inv = 1.f / (kPi * base * base * base);
    g_a2 += gD * (1.f - (a2 + 1.f) * c2) * inv;
    if (n_h > kEps && n_h < 1.f - kEps) g_n_h += gD * -(4.f * (a2 - 1.f) * a2 * n_h) * inv;
  }
  V3 g_h = s.n * g_n_h + s.wo * g_wo_h;
  out.nrm += h * g_n_h + wi * g_wi_n + s.wo * g_wo_n;
  V3 g_wo = h * g_wo_h + s.n * g_wo_n + normalize0_bwd(h_raw, g_h);
  float g_alpha = (s.alpha > kMinRough * kMinRough) ? g_a2 * 2.f * s.alpha : 0.f;
  // spec_col = (0.04 (1 - m) + kd m)(1 - x),  alpha = y^2,  wo = normalize(view - pos)
  float x = s.arm.x, m = s.arm.z;
  out.kd += g_col * ((1.f - x) * m);
  out.arm.x += dot(g_col, (v3(0.04f) - s.kd) * m - v3(0.04f));
  out.arm.z += dot(g_col, s.kd - v3(0.04f)) * (1.f - x);
  out.arm.y += g_alpha * 2.f * s.arm.y;
  out.pos -= normalize0_bwd(s.wo_raw, g_wo);
}

__device__ __forceinline__ int cdf_steps(int n) { return (int)ceilf(log2f((float)(n - 1))) + 1; }

// PCG's LCG jumped ahead by n steps (O(log n)); lets a launch start at any sample pair
__device__ __forceinline__ uint32_t lcg_skip(uint32_t state, uint32_t n) {
  uint32_t cur_mult = 747796405u, cur_plus = 2891336453u, acc_mult = 1u, acc_plus = 0u;
  while (n) {
    if (n & 1u) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
    cur_plus = (cur_mult + 1u) * cur_plus;
    cur_mult *= cur_mult;
    n >>= 1;
  }
  return acc_mult * state + acc_plus;
}

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_GEN = 2 };
// resident CTAs of 128 threads per SM that the register allocation of each mode must allow (BWD 128 registers, FWD / GEN 72)
#ifndef GSB_SHADE_BWD_BLOCKS
#define GSB_SHADE_BWD_BLOCKS 4
#endif
#ifndef GSB_SHADE_FWD_BLOCKS
#define GSB_SHADE_FWD_BLOCKS 7
#endif
#ifndef GSB_SHADE_GEN_BLOCKS
#define GSB_SHADE_GEN_BLOCKS 7
#endif
#ifndef GSB_GEN_PAIRS
#define GSB_GEN_PAIRS 4              // sample pairs whose rays one lane gathers before the warp appends them
#endif

// One thread per pixel, sample pairs [i0, i1).  MODE_GEN only regenerates the sample directions and stores those that
// need a shadow ray; the rays are traced by k_trace_pool (occluder.cu) with repacked lanes, and MODE_FWD / MODE_BWD
// consume the resulting visibility.  (Tracing inline made the warp wait for its slowest ray on every sample: ncu showed
// 2.4 of 32 lanes active.)
template <int MODE>
__global__ void __launch_bounds__(128, MODE == 1 ? GSB_SHADE_BWD_BLOCKS : (MODE == 0 ? GSB_SHADE_FWD_BLOCKS : GSB_SHADE_GEN_BLOCKS)) k_env_shade(ShadeParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int b = blockIdx.z;
  if (x >= p.W || y >= p.H) return;
  const size_t npix = (size_t)p.B * p.H * p.W;
  const size_t pix = ((size_t)b * p.H + y) * p.W + x;
  if (!(__ldg(p.mask + pix) > 0.f)) {        // masked pixels: outputs stay zero (kernel.cu:478)
    if (MODE == MODE_BWD && p.first_chunk) {
      st3(p.g_pos + pix * 3, v3(0.f)); st3(p.g_nrm + pix * 3, v3(0.f));
      st3(p.g_kd + pix * 3, v3(0.f)); st3(p.g_ks + pix * 3, v3(0.f));
    } else if (MODE == MODE_FWD && p.first_chunk) {
      st3(p.diff + pix * 3, v3(0.f)); st3(p.spec + pix * 3, v3(0.f));
    }
    return;
  }
  SurfaceConst s;
  const V3 pos = ld3(p.pos + pix * 3), vpos = ld3(p.view_pos + (size_t)b * 3);
  V3 origin = v3(0.f);
  if (MODE == MODE_GEN) origin = ld3(p.ro + pix * 3);
  s.n = ld3(p.nrm + pix * 3);
  s.kd = ld3(p.kd + pix * 3);
  s.arm = ld3(p.ks + pix * 3);
  s.wo_raw = vpos - pos;
  s.wo = normalize0(s.wo_raw);
  s.alpha = s.arm.y * s.arm.y;
  s.spec_col = (v3(0.04f) * (1.f - s.arm.z) + s.kd * s.arm.z) * (1.f - s.arm.x);
  const Frame frame = make_frame(s.n);

  // lobe selection probabilities (kernel.cu:495-502, albedo :81-94)
  const float metallic = s.arm.z;
  const V3 f0 = v3(0.04f) * (1.f - metallic) + s.kd * metallic;
  const float w_d = (1.f - metallic) * luminance(s.kd);
  float w_s = 0.f;
  {
    float cos_no = normalize0(to_local(frame, s.wo)).z;
    if (cos_no > 0.f) {
      float sc = pow5(1.f - clampf(cos_no, kEps, 1.f - kEps));
      w_s = luminance(f0 * (1.f - sc) + v3(sc));
    }
  }
  const float p_d = (w_d + w_s) > 0.f ? w_d / (w_d + w_s) : 1.f;
  const float p_s = 1.f - p_d;

  uint32_t rng = pcg_hash(p.seed, p.pixel_ids ? __ldg(p.pixel_ids + pix) : (uint32_t)pix);
  const uint32_t light_row = pcg_next(rng) % (uint32_t)p.n_perms;
  const uint32_t bsdf_row = pcg_next(rng) % (uint32_t)p.n_perms;
  if (p.i0 > 0) rng = lcg_skip(rng, 5u * (uint32_t)p.i0);           // 5 uniforms per sample pair
  const int n = p.n, n2 = n * n;
  const int32_t* __restrict__ perm_l = p.perms + (size_t)light_row * n2;
  const int32_t* __restrict__ perm_b = p.perms + (size_t)bsdf_row * n2;
  const float strata = 1.0f / (float)n, weight = 1.0f / (float)n2;
  const uint32_t n_magic = (uint32_t)((0x100000000ull + (uint64_t)n - 1) / (uint64_t)n);   // exact st / n for st < 2^32 / n
  const int steps_r = cdf_steps(p.lh), steps_c = cdf_steps(p.lw);

  V3 gd = v3(0.f), gs = v3(0.f);
  if (MODE == MODE_BWD) {
    gd = ld3(p.g_diff + pix * 3);
    gs = ld3(p.g_spec + pix * 3);
  }
  V3 acc_d = v3(0.f), acc_s = v3(0.f);
  PixelGrads pg;
  pg.kd = pg.arm = pg.pos = pg.nrm = v3(0.f);
  uint32_t vis_word = 0xffffffffu;     // bit k of word w = sample 32 w + k visible (samples in process() order)
  int sample_id = 2 * p.i0;            // global sample index (2 per pair)
  int local_id = 0;                    // index inside this chunk

  auto process = [&](V3 dir, int tex, float pdf_sum) {              // process_sample (kernel.cu:403-461)
    const V3 L = ld3(p.light + (size_t)tex * 3);
    const float mis = 1.0f / fmaxf(pdf_sum, 0.0001f);
    float fd;
    V3 fs;
    eval_bsdf(s, dir, p.bsdf, fd, fs);
    // shadow term (kernel.cu:101-118, :420): V = vis*s + (1-s), vis from the traced chunk or the forward's bit record
    float vis = 1.0f;
    if (MODE == MODE_BWD && p.vis_in) {
      if ((sample_id & 31) == 0 || local_id == 0) vis_word = __ldg(p.vis_in + pix * p.vis_words + (sample_id >> 5));
      vis = (vis_word >> (sample_id & 31)) & 1u ? 1.f : 0.f;
    } else if (p.vis_chunk) {
      vis = p.vis_chunk[(size_t)local_id * npix + pix] ? 1.f : 0.f;
    }
    if (MODE == MODE_FWD && p.vis_out) {
      if (vis == 0.f) vis_word &= ~(1u << (sample_id & 31));
      if ((sample_id & 31) == 31) {
        p.vis_out[pix * p.vis_words + (sample_id >> 5)] = vis_word;
        vis_word = 0xffffffffu;
      }
    }
    ++sample_id;
    ++local_id;
    const float V = vis * p.shadow_scale + (1.f - p.shadow_scale);
    const float k = V * mis * weight;
    if (MODE == MODE_FWD) {
      acc_d += L * (fd * k);
      acc_s += fs * L * k;
    } else {
      V3 gl = (gd * fd + gs * fs) * k;
      // one 16-byte vector reduction per sample into the padded [lh, lw, 4] gradient probe (three scalar atomics before)
      float* t = p.g_light + (size_t)tex * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t), "f"(gl.x), "f"(gl.y), "f"(gl.z), "f"(0.f) : "memory");
      eval_bsdf_bwd(s, dir, p.bsdf, dot(gd, L) * k, gs * L * k, pg);
    }
  };

  // the two stratified samples of pair i (kernel.cu:463-474): 2 uniforms for the light importance sample, then 3 for the BSDF one
  auto sample_light = [&](int i, float& pdf_l, int& tex_l) {
    const int st = __ldg(perm_l + i);
    const int sq = (int)__umulhi((uint32_t)st, n_magic);
    const float sx = ((float)(st - sq * n) + pcg_uniform(rng)) * strata;
    const float sy = ((float)sq + pcg_uniform(rng)) * strata;
    return light_sample(p, steps_r, steps_c, sx, sy, pdf_l, tex_l);
  };
  auto sample_bsdf = [&](int i, float& pdf_b) {
    const int st = __ldg(perm_b + i);
    const int sq = (int)__umulhi((uint32_t)st, n_magic);
    const float sx = ((float)(st - sq * n) + pcg_uniform(rng)) * strata;
    const float sy = ((float)sq + pcg_uniform(rng)) * strata;
    const float sz = pcg_uniform(rng);
    return bsdf_sample(frame, p_d, p_s, s.n, s.wo, sx, sy, sz, s.alpha, pdf_b);
  };

  if (MODE == MODE_GEN) {
    // Shadow rays of kGenPairs sample pairs are gathered in registers and appended with ONE counter update per warp (an
    // exclusive scan over the lanes' ray counts): the single list counter was the limiter of this mode (ncu r2f: 8.4 M
    // same-address atomics per launch, 4 of 7.6 ms).  A sample can only contribute if it lies in the upper hemisphere of the
    // shading normal (Lambert > 0; the GGX lobe additionally needs n.wi > 1e-4): everything else gets no shadow ray -- the
    // reference traces those too, and then multiplies their visibility by a zero BSDF value.
    namespace cg = cooperative_groups;
    constexpr int kGenPairs = GSB_GEN_PAIRS;
    cg::coalesced_group grp = cg::coalesced_threads();             // the unmasked pixels of this warp; stable over the loop
    for (int i = p.i0; i < p.i1; i += kGenPairs) {
      V3 dirs[2 * kGenPairs];
      uint32_t want = 0u;
#pragma unroll
      for (int j = 0; j < kGenPairs; ++j) {
        dirs[2 * j] = dirs[2 * j + 1] = v3(0.f);
        if (i + j < p.i1) {
          float pdf_l, pdf_b;
          int tex;
          dirs[2 * j] = sample_light(i + j, pdf_l, tex);
          dirs[2 * j + 1] = sample_bsdf(i + j, pdf_b);
          if (dot(s.n, dirs[2 * j]) > 0.f) want |= 1u << (2 * j);
          if (dot(s.n, dirs[2 * j + 1]) > 0.f) want |= 2u << (2 * j);
        }
      }
      const int cnt = __popc(want);
      const int before = cg::exclusive_scan(grp, cnt);
      const int total = grp.shfl(before + cnt, grp.size() - 1);
      int base = 0;
      if (grp.thread_rank() == 0 && total > 0) base = atomicAdd(p.ray_count, total);
      int slot = grp.shfl(base, 0) + before;
#pragma unroll
      for (int q = 0; q < 2 * kGenPairs; ++q) {
        if ((want >> q) & 1u) {
          const int rid = (int)((size_t)(local_id + q) * npix + pix);       // index into this chunk's visibility bytes
          if (slot < p.ray_cap) {     // never false when the caller's n_covered is a true upper bound of the unmasked pixels
            const size_t e = 2 * (size_t)slot;
            p.ray_list[e] = make_float4(origin.x, origin.y, origin.z, __int_as_float(rid));
            p.ray_list[e + 1] = make_float4(dirs[q].x, dirs[q].y, dirs[q].z, 0.f);
          } else {
            atomicAdd(p.dropped, 1u);   // reported as an error by the next gsb_env_shade_* call / gsb_env_shade_dropped_rays()
          }
          ++slot;
        }
      }
      local_id += 2 * kGenPairs;
    }
  } else {
  for (int i = p.i0; i < p.i1; ++i) {
    float pdf_light, pdf_b;
    int tex;
    // (1) light importance sample
    V3 dir = sample_light(i, pdf_light, tex);
    process(dir, tex, pdf_light + bsdf_pdf(frame, p_d, p_s, s.n, s.wo, dir, s.alpha));
    // (2) BSDF importance sample
    dir = sample_bsdf(i, pdf_b);
    tex = 0;
    pdf_light = light_pdf(p, dir, tex);
    process(dir, tex, pdf_light + pdf_b);
  }
  }
  if (MODE == MODE_FWD) {
    if (p.vis_out && (sample_id & 31) != 0) p.vis_out[pix * p.vis_words + (sample_id >> 5)] = vis_word;
    if (!p.first_chunk) { acc_d += ld3(p.diff + pix * 3); acc_s += ld3(p.spec + pix * 3); }
    st3(p.diff + pix * 3, acc_d);
    st3(p.spec + pix * 3, acc_s);
  } else if (MODE == MODE_BWD) {
    if (!p.first_chunk) {
      pg.pos += ld3(p.g_pos + pix * 3); pg.nrm += ld3(p.g_nrm + pix * 3);
      pg.kd += ld3(p.g_kd + pix * 3); pg.arm += ld3(p.g_ks + pix * 3);
    }
    st3(p.g_pos + pix * 3, pg.pos);
    st3(p.g_nrm + pix * 3, pg.nrm);
    st3(p.g_kd + pix * 3, pg.kd);
    st3(p.g_ks + pix * 3, pg.arm);
  }
}

int fill(ShadeParams& p, const float* mask, const float* ro, const float* pos, const float* nrm, const float* view_pos,
         const float* kd, const float* ks, const float* light, const float* pdf, const float* rows, const float* cols,
         const float* rows_top, const float* cols_top, const int32_t* perms, int64_t B, int64_t H, int64_t W, int64_t lh, int64_t lw, int64_t n_perms, int bsdf,
         int n_samples_x, uint32_t seed, float shadow_scale) {
  if (bsdf < 0 || bsdf > 2 || n_samples_x < 1 || lh < 2 || lw < 2 || n_perms < 1) return (int)cudaErrorInvalidValue;
  p.mask = mask; p.ro = ro; p.pos = pos; p.nrm = nrm; p.view_pos = view_pos; p.kd = kd; p.ks = ks;
  p.light = light; p.pdf = pdf; p.rows = rows; p.cols = cols; p.perms = perms;
  const bool aux_ok = rows_top && cols_top && lh % 16 == 0 && lw % 16 == 0 && lh <= 256 && lw <= 256;
  p.rows_top = aux_ok ? rows_top : nullptr;
  p.cols_top = aux_ok ? cols_top : nullptr;
  p.B = (int)B; p.H = (int)H; p.W = (int)W; p.lh = (int)lh; p.lw = (int)lw; p.n_perms = (int)n_perms;
  p.bsdf = bsdf; p.n = n_samples_x; p.seed = seed; p.shadow_scale = shadow_scale;
  p.g_diff = p.g_spec = nullptr;
  p.ray_list = nullptr; p.ray_count = nullptr; p.vis_chunk = nullptr; p.dropped = nullptr; p.pixel_ids = nullptr;
  p.vis_out = nullptr; p.vis_in = nullptr; p.vis_words = (2 * n_samples_x * n_samples_x + 31) / 32;
  p.i0 = 0; p.i1 = n_samples_x * n_samples_x; p.first_chunk = 1;
  p.diff = p.spec = p.g_pos = p.g_nrm = p.g_kd = p.g_ks = p.g_light = nullptr;
  return 0;
}

}  // namespace

// trace kernel lives in occluder.cu
extern "C" int gsb_trace_shadow_rays(const void* occluder, const void* ray_list, const int32_t* ray_count, int64_t ray_cap,
                                     int32_t* fetch_counter, uint8_t* vis, void* stream);

namespace {

// scratch per sample pair: worst-case ray list (2 rays per UNMASKED pixel x 32 B) + visibility bytes (dense over all pixels);
// plus 256 B of counters
inline int64_t covered_bound(int64_t npix, int64_t n_covered) { return (n_covered <= 0 || n_covered > npix) ? npix : n_covered; }
inline size_t pair_bytes(int64_t npix, int64_t ncov) { return (size_t)ncov * 2 * 2 * sizeof(float4) + (size_t)npix * 2; }
constexpr size_t kCounterBytes = 256;
inline int pairs_per_chunk(int64_t npix, int64_t ncov, int n2, size_t scratch_bytes) {
  if (scratch_bytes <= kCounterBytes) return 0;
  int64_t ppc = (int64_t)((scratch_bytes - kCounterBytes) / pair_bytes(npix, ncov));
  const int64_t id_limit = ((int64_t)1 << 31) / (2 * npix);    // ray ids (2 * pairs * npix) and the list counter are int32
  if (ppc > id_limit) ppc = id_limit;
  if (ppc >= n2) return n2;
  ppc = ppc / 16 * 16;                 // chunk borders on 32-sample words of the visibility bit record
  return (int)ppc;
}

// optional device timing of the trace launches (bench.py's roofline leg): events around every gsb_trace_shadow_rays call
constexpr int kTimedLaunches = 1024;
struct TraceTimer {
  bool enabled = false;
  int used = 0;
  int total = 0;                    // trace launches since timing was enabled (events exist for the first kTimedLaunches)
  int64_t rays_pixels = 0;          // sum over chunks of n_pix * layers (upper bound of rays, masked/unlit included)
  cudaEvent_t ev[2 * kTimedLaunches];
  bool created = false;
} g_timer;

// Rays that did not fit the list of their chunk (only possible when the caller understated n_covered).  The device counter is
// mirrored into pinned host memory by an async copy at the end of every traced call, so the NEXT call can refuse without a sync.
struct DropState {
  unsigned int* d_count = nullptr;
  volatile unsigned int* h_count = nullptr;
  bool ready = false;
} g_drop;
bool drop_init() {
  if (g_drop.ready) return true;
  if (cudaMalloc(&g_drop.d_count, sizeof(unsigned int)) != cudaSuccess) return false;
  if (cudaMemset(g_drop.d_count, 0, sizeof(unsigned int)) != cudaSuccess) return false;
  unsigned int* h = nullptr;
  if (cudaMallocHost(&h, sizeof(unsigned int)) != cudaSuccess) return false;
  *h = 0u;
  g_drop.h_count = h;
  g_drop.ready = true;
  return true;
}

template <int MODE>
void launch(const ShadeParams& p, cudaStream_t stream) {
  dim3 block(16, 8), grid((unsigned)((p.W + 15) / 16), (unsigned)((p.H + 7) / 8), (unsigned)p.B);
  k_env_shade<MODE><<<grid, block, 0, stream>>>(p);
}

// runs MODE over all sample pairs, tracing shadow rays chunk by chunk when an occluder is given
template <int MODE>
int run(ShadeParams p, const void* bvh, void* scratch, size_t scratch_bytes, int64_t n_covered, cudaStream_t stream) {
  const int n2 = p.n * p.n;
  const int64_t npix = (int64_t)p.B * p.H * p.W;
  const bool replay = MODE == MODE_BWD && p.vis_in != nullptr;
  if (!bvh || !(p.shadow_scale > 0.f) || replay) {
    launch<MODE>(p, stream);
    return (int)cudaGetLastError();
  }
  const int64_t ncov = covered_bound(npix, n_covered);
  const int ppc = scratch ? pairs_per_chunk(npix, ncov, n2, scratch_bytes) : 0;
  if (ppc < 1) return (int)cudaErrorInvalidValue;       // shadow rays need scratch for at least 16 sample pairs
  if (!drop_init()) return (int)cudaErrorMemoryAllocation;
  if (*g_drop.h_count != 0u) return (int)cudaErrorInvalidValue;      // an earlier call lost rays: its n_covered was too small
  int* counters = (int*)scratch;                                       // [0] list length, [1] trace fetch cursor
  float4* list = (float4*)((char*)scratch + kCounterBytes);
  const int64_t cap = ncov * 2 * ppc;
  uint8_t* vis = (uint8_t*)scratch + kCounterBytes + (size_t)cap * 2 * sizeof(float4);
  for (int i0 = 0; i0 < n2; i0 += ppc) {
    ShadeParams q = p;
    q.i0 = i0;
    q.i1 = i0 + ppc < n2 ? i0 + ppc : n2;
    q.first_chunk = i0 == 0;
    q.ray_list = list;
    q.ray_count = counters;
    q.ray_cap = (int)(cap < 0x7fffffff ? cap : 0x7fffffff);
    q.dropped = g_drop.d_count;
    cudaError_t e = cudaMemsetAsync(counters, 0, kCounterBytes, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(vis, 1, (size_t)npix * 2 * (q.i1 - q.i0), stream);   // everything visible until hit
    if (e != cudaSuccess) return (int)e;
    launch<MODE_GEN>(q, stream);
    if (g_timer.enabled) ++g_timer.total;
    const bool timed = g_timer.enabled && g_timer.used < kTimedLaunches;
    if (timed) cudaEventRecord(g_timer.ev[2 * g_timer.used], stream);
    int err = gsb_trace_shadow_rays(bvh, list, counters, cap, counters + 1, vis, (void*)stream);
    if (timed) {
      cudaEventRecord(g_timer.ev[2 * g_timer.used + 1], stream);
      ++g_timer.used;
      g_timer.rays_pixels += npix * 2 * (int64_t)(q.i1 - q.i0);
    }
    if (err) return err;
    q.ray_list = nullptr;
    q.vis_chunk = vis;
    launch<MODE>(q, stream);
  }
  cudaMemcpyAsync((void*)g_drop.h_count, g_drop.d_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

/* Profiling aid: gsb_trace_timing(1) starts recording CUDA events around the trace launches of subsequent env_shade calls
 * (up to 1024); gsb_trace_timing(0) stops and returns the summed device time in milliseconds (synchronises). */
float gsb_trace_timing(int enable) {
  if (!g_timer.created) {
    for (int i = 0; i < 2 * kTimedLaunches; ++i) cudaEventCreate(&g_timer.ev[i]);
    g_timer.created = true;
  }
  if (enable) {
    g_timer.enabled = true;
    g_timer.used = 0;
    g_timer.total = 0;
    g_timer.rays_pixels = 0;
    return 0.f;
  }
  g_timer.enabled = false;
  float total = 0.f;
  for (int i = 0; i < g_timer.used; ++i) {
    cudaEventSynchronize(g_timer.ev[2 * i + 1]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_timer.ev[2 * i], g_timer.ev[2 * i + 1]);
    total += ms;
  }
  return total;
}

/* Shadow rays lost because a chunk's ray list was full, i.e. the caller's n_covered was below the number of pixels with
 * mask > 0 (synchronises the device; reset != 0 clears the counter).  While the count is non-zero every traced
 * gsb_env_shade_* call fails with cudaErrorInvalidValue: results computed with lost rays are wrong, never silently so. */
uint32_t gsb_env_shade_dropped_rays(int reset) {
  if (!drop_init()) return 0xffffffffu;
  cudaDeviceSynchronize();
  unsigned int v = 0u;
  cudaMemcpy(&v, g_drop.d_count, sizeof(v), cudaMemcpyDeviceToHost);
  if (reset) {
    cudaMemset(g_drop.d_count, 0, sizeof(unsigned int));
    *g_drop.h_count = 0u;
  } else {
    *g_drop.h_count = v;
  }
  return v;
}

/* Trace launches since the last gsb_trace_timing(1); the summed time covers the first 1024 of them. */
int gsb_trace_launches(void) { return g_timer.total; }

size_t gsb_env_shade_scratch_bytes(int64_t B, int64_t H, int64_t W, int64_t n_covered, int n_samples_x, size_t budget_bytes) {
  const int64_t npix = B * H * W;
  const int n2 = n_samples_x * n_samples_x;
  if (npix == 0) return 0;
  const int64_t ncov = covered_bound(npix, n_covered);
  int ppc = pairs_per_chunk(npix, ncov, n2, budget_bytes);
  if (ppc < 16) ppc = 16 < n2 ? 16 : n2;
  return pair_bytes(npix, ncov) * (size_t)ppc + kCounterBytes;
}

/* Chunks of sample pairs a traced env_shade call is split into for this scratch size (3 kernels per chunk). */
int gsb_env_shade_chunks(int64_t B, int64_t H, int64_t W, int64_t n_covered, int n_samples_x, size_t scratch_bytes) {
  const int64_t npix = B * H * W;
  const int n2 = n_samples_x * n_samples_x;
  if (npix == 0) return 0;
  const int ppc = pairs_per_chunk(npix, covered_bound(npix, n_covered), n2, scratch_bytes);
  return ppc < 1 ? 0 : (n2 + ppc - 1) / ppc;
}

int gsb_env_shade_fwd(const float* mask, const float* ro, const float* pos, const float* nrm, const float* view_pos,
                      const float* kd, const float* ks, const float* light, const float* pdf, const float* rows,
                      const float* cols, const float* rows_top, const float* cols_top, const int32_t* perms, int64_t B, int64_t H,
                      int64_t W, int64_t lh, int64_t lw, int64_t n_perms, int bsdf, int n_samples_x, uint32_t rnd_seed,
                      float shadow_scale, const void* bvh, void* scratch, size_t scratch_bytes, int64_t n_covered, uint32_t* vis_bits, float* diff,
                      float* spec, const uint32_t* pixel_ids, void* stream) {
  ShadeParams p;
  int err = fill(p, mask, ro, pos, nrm, view_pos, kd, ks, light, pdf, rows, cols, rows_top, cols_top, perms, B, H, W, lh, lw, n_perms, bsdf,
                 n_samples_x, rnd_seed, shadow_scale);
  if (err) return err;
  if (B * H * W == 0) return 0;
  p.diff = diff; p.spec = spec;
  p.pixel_ids = pixel_ids;
  p.vis_out = (bvh && shadow_scale > 0.f) ? vis_bits : nullptr;
  return run<MODE_FWD>(p, bvh, scratch, scratch_bytes, n_covered, (cudaStream_t)stream);
}

int gsb_env_shade_bwd(const float* mask, const float* ro, const float* pos, const float* nrm, const float* view_pos,
                      const float* kd, const float* ks, const float* light, const float* pdf, const float* rows,
                      const float* cols, const float* rows_top, const float* cols_top, const int32_t* perms, int64_t B, int64_t H,
                      int64_t W, int64_t lh, int64_t lw, int64_t n_perms, int bsdf, int n_samples_x, uint32_t rnd_seed,
                      float shadow_scale, const void* bvh, void* scratch, size_t scratch_bytes, int64_t n_covered, const uint32_t* vis_bits,
                      const float* g_diff, const float* g_spec, float* g_pos, float* g_nrm, float* g_kd, float* g_ks,
                      float* g_light, const uint32_t* pixel_ids, void* stream) {
  ShadeParams p;
  int err = fill(p, mask, ro, pos, nrm, view_pos, kd, ks, light, pdf, rows, cols, rows_top, cols_top, perms, B, H, W, lh, lw, n_perms, bsdf,
                 n_samples_x, rnd_seed, shadow_scale);
  if (err) return err;
  cudaError_t e = cudaMemsetAsync(g_light, 0, sizeof(float) * 4 * (size_t)lh * lw, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  if (B * H * W == 0) return 0;
  p.g_diff = g_diff; p.g_spec = g_spec;
  p.pixel_ids = pixel_ids;
  p.g_pos = g_pos; p.g_nrm = g_nrm; p.g_kd = g_kd; p.g_ks = g_ks; p.g_light = g_light;
  p.vis_in = (bvh && shadow_scale > 0.f) ? vis_bits : nullptr;
  return run<MODE_BWD>(p, bvh, scratch, scratch_bytes, n_covered, (cudaStream_t)stream);
}

}  // extern "C"
