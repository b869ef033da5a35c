// Tangent frame of the extracted mesh, forward and adjoint, for sm_100a.
//
// Replaces the reference's map_uv / compute_tangents / boundary extension (geometry/gshell_tets.py:40-78, 210-239, 318-319,
// 337-338, 375-380): a 4 * ceil(sqrt(T))^2-row uv table (415 MB at the "256" grid), six scatter_add_ passes, two normalisations
// and two gathers-with-weights, ~40 ATen kernels plus their autograd twins -- with 3 + 4 kernels and no table: the uv of atlas row
// r is arithmetic on r (cell r / 4 of an N x N grid, corner r % 4).
//
// Reference quirk kept on purpose: compute_tangents is called with the VERTEX ids of a face as its uv indices (:319), so the
// "uv" of vertex v is atlas row v.
//
// HBM / atomic bound: 3 x 36 B gathered (L2) + 12 red.global.add per face, 28 B per vertex, 2 x 16 B gathered per boundary vertex.
// Compiled with -fmad=false: the per-face tangent reproduces the separately rounded mul / sub / div of the PyTorch ops it
// replaces (the sums over a vertex's faces still differ in order -- atomics).
// One independent thread per element in every kernel: this unit also compiles as plain host code for the CPU tests
// (tests/native/host_kernels.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"
#include "vec.cuh"

using namespace gsb;

namespace {
constexpr int kThreads = 256;
constexpr float kEps = 1e-20f;                     // clamp of the squared length in both normalisations
inline int nblk(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

struct Atlas {
  const float* lin;                                // [n] = linspace(0, 1 - 1/n, n), built by the host with the reference's op
  int n;
  float pad;                                       // 0.9 / n
};

struct UV { float x, y; };
// row r of the atlas map_uv builds (:210-225): the 4 corners of cell r / 4
__device__ __forceinline__ UV atlas_uv(const Atlas& a, int r) {
  const int cell = r >> 2, corner = r & 3;
  const int cx = cell % a.n, cy = min(cell / a.n, a.n - 1);      // a row past the table is an error in the reference; stay in bounds
  UV t{__ldg(a.lin + cx), __ldg(a.lin + cy)};
  if (corner == 1 || corner == 2) t.x = t.x + a.pad;
  if (corner >= 2) t.y = t.y + a.pad;
  return t;
}

// per-face tangent pieces shared by the forward and the adjoint
struct FaceFrame { float du1y, du2y, denc; };
__device__ __forceinline__ FaceFrame face_frame(const Atlas& a, int i0, int i1, int i2) {
  const UV t0 = atlas_uv(a, i0), t1 = atlas_uv(a, i1), t2 = atlas_uv(a, i2);
  const float du1x = t1.x - t0.x, du1y = t1.y - t0.y, du2x = t2.x - t0.x, du2y = t2.y - t0.y;
  const float den = du1x * du2y - du1y * du2x;
  return FaceFrame{du1y, du2y, den > 0.f ? fmaxf(den, 1e-6f) : fminf(den, -1e-6f)};
}

// acc[v] += (tangent of the face, 1) for one corner of one face (compute_tangents :52-70).  Item i = corner * F + face: the
// launch order is the order of the reference's three scatter_add_ passes (corner 0 of every face, then corner 1, then corner 2),
// so that -- as far as atomics in launch order go -- a vertex sums its faces in the reference's order.  That matters only for
// vertices whose face tangents cancel: their normalised sum is rounding noise in ANY implementation
// (tests/test_oracle_tangent_conditioning.py), but with the same order it tends to be the same noise.
__global__ void __launch_bounds__(kThreads) k_tng_face(const float* __restrict__ v, const int32_t* __restrict__ tris, int64_t n_faces,
                                                       Atlas atlas, float* __restrict__ acc) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= 3 * n_faces) return;
  const int corner = (int)(i / n_faces);
  const int64_t f = i - corner * n_faces;
  const int idx[3] = {__ldg(tris + f * 3), __ldg(tris + f * 3 + 1), __ldg(tris + f * 3 + 2)};
  const V3 p0 = ld3(v + (size_t)idx[0] * 3), p1 = ld3(v + (size_t)idx[1] * 3), p2 = ld3(v + (size_t)idx[2] * 3);
  const FaceFrame ff = face_frame(atlas, idx[0], idx[1], idx[2]);
  const V3 tang = ((p1 - p0) * ff.du2y - (p2 - p0) * ff.du1y) / ff.denc;
  float* a = acc + (size_t)idx[corner] * 4;
  atomicAdd(a, tang.x); atomicAdd(a + 1, tang.y); atomicAdd(a + 2, tang.z); atomicAdd(a + 3, 1.f);
}

// x / sqrt(max(|x|^2, eps)) and its adjoint (the clamp's branch passes no gradient to |x|^2)
__device__ __forceinline__ V3 unit(V3 x) { return x / sqrtf(fmaxf(dot(x, x), kEps)); }
__device__ __forceinline__ V3 unit_bwd(V3 x, V3 g) {
  const float s = dot(x, x);
  if (!(s > kEps)) return g / sqrtf(kEps);
  const float inv = 1.f / sqrtf(s);
  const V3 xh = x * inv;
  return (g - xh * dot(xh, g)) * inv;
}

// mean tangent -> unit -> Gram-Schmidt against the vertex normal -> unit (:72-77)
__global__ void __launch_bounds__(kThreads) k_tng_vertex(const float* __restrict__ acc, const float* __restrict__ nrm, int64_t n_verts,
                                                         float* __restrict__ tng) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_verts) return;
  const float c = __ldg(acc + i * 4 + 3);
  const V3 u = unit(ld3(acc + i * 4) / c);
  const V3 n = ld3(nrm + i * 3);
  st3(tng + i * 3, unit(u - n * dot(u, n)));
}

__global__ void __launch_bounds__(kThreads) k_tng_vertex_bwd(const float* __restrict__ acc, const float* __restrict__ nrm,
                                                             int64_t n_verts, float* __restrict__ g_t /* in: d/d tng, out: d/d acc.xyz */,
                                                             float* __restrict__ g_nrm) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_verts) return;
  const float c = __ldg(acc + i * 4 + 3);
  const V3 t0 = ld3(acc + i * 4) / c;
  const V3 u = unit(t0);
  const V3 n = ld3(nrm + i * 3);
  const float d = dot(u, n);
  const V3 p = u - n * d;
  const V3 g_p = unit_bwd(p, V3{g_t[i * 3], g_t[i * 3 + 1], g_t[i * 3 + 2]});
  const float ng = dot(n, g_p);
  st3(g_nrm + i * 3, -(g_p * d + u * ng));
  st3(g_t + i * 3, unit_bwd(t0, g_p - n * ng) / c);
}

__global__ void __launch_bounds__(kThreads) k_tng_face_bwd(const int32_t* __restrict__ tris, int64_t n_faces, Atlas atlas,
                                                           const float* __restrict__ g_acc, float* __restrict__ g_v) {
  const int64_t f = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (f >= n_faces) return;
  const int i0 = __ldg(tris + f * 3), i1 = __ldg(tris + f * 3 + 1), i2 = __ldg(tris + f * 3 + 2);
  const FaceFrame ff = face_frame(atlas, i0, i1, i2);
  const V3 g_nom = (ld3(g_acc + (size_t)i0 * 3) + ld3(g_acc + (size_t)i1 * 3) + ld3(g_acc + (size_t)i2 * 3)) / ff.denc;
  const V3 g1 = g_nom * ff.du2y, g2 = -(g_nom * ff.du1y), g0 = -(g1 + g2);
  float* o0 = g_v + (size_t)i0 * 3;
  float* o1 = g_v + (size_t)i1 * 3;
  float* o2 = g_v + (size_t)i2 * 3;
  atomicAdd(o0, g0.x); atomicAdd(o0 + 1, g0.y); atomicAdd(o0 + 2, g0.z);
  atomicAdd(o1, g1.x); atomicAdd(o1 + 1, g1.y); atomicAdd(o1 + 2, g1.z);
  atomicAdd(o2, g2.x); atomicAdd(o2 + 1, g2.y); atomicAdd(o2 + 2, g2.z);
}

// Boundary vertex j sits on the polygon edge (a, b): a = its slot's vertex, b = the next vertex of the same polygon (3 slots
// per triangle polygon first, then 4 per quad polygon).  Weights = the mSDF zero crossing on that edge, zero where the edge
// does not straddle (:345-365); the positions use the same weights inside the extraction.
struct Boundary { int a, b; float w0, w1, ma, mb; bool ok; };
__device__ __forceinline__ Boundary boundary(const int32_t* __restrict__ slot_a, const float* __restrict__ msdf, int64_t j, int64_t n3) {
  int64_t base, next;
  if (j < n3) { base = j - j % 3; next = base + (j - base + 1) % 3; }
  else { const int64_t q = j - n3; base = n3 + (q / 4) * 4; next = base + (q % 4 + 1) % 4; }
  Boundary r;
  r.a = __ldg(slot_a + j) & 0x7FFFFFFF;
  r.b = __ldg(slot_a + next) & 0x7FFFFFFF;
  r.ma = __ldg(msdf + r.a);
  r.mb = __ldg(msdf + r.b);
  const int sa = (r.ma > 0.f) - (r.ma < 0.f), sb = (r.mb > 0.f) - (r.mb < 0.f);
  const float den = r.ma - r.mb;
  r.ok = abs(sa + sb) != 2 && fabsf(den) > 1e-12f;
  const float safe = r.ok ? den : 1.f;
  r.w0 = r.ok ? -r.mb / safe : 0.f;
  r.w1 = r.ok ? r.ma / safe : 0.f;
  return r;
}

__global__ void __launch_bounds__(kThreads) k_tng_boundary(const int32_t* __restrict__ slot_a, const float* __restrict__ msdf,
                                                           int64_t n_boundary, int64_t n3, int64_t n_wt, float* __restrict__ tng) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= n_boundary) return;
  const Boundary e = boundary(slot_a, msdf, j, n3);
  const V3 ta = V3{tng[(size_t)e.a * 3], tng[(size_t)e.a * 3 + 1], tng[(size_t)e.a * 3 + 2]};
  const V3 tb = V3{tng[(size_t)e.b * 3], tng[(size_t)e.b * 3 + 1], tng[(size_t)e.b * 3 + 2]};
  st3(tng + (size_t)(n_wt + j) * 3, ta * e.w0 + tb * e.w1);
}

// g_t starts as d/d tng of the watertight rows; the boundary rows add theirs (and the gradient of the weights to the mSDF)
__global__ void __launch_bounds__(kThreads) k_tng_grad_init(const float* __restrict__ g_tng, int64_t n_wt, float* __restrict__ g_t,
                                                            float* __restrict__ g_msdf) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_wt) return;
  g_t[i * 3] = __ldg(g_tng + i * 3); g_t[i * 3 + 1] = __ldg(g_tng + i * 3 + 1); g_t[i * 3 + 2] = __ldg(g_tng + i * 3 + 2);
  g_msdf[i] = 0.f;
}

__global__ void __launch_bounds__(kThreads) k_tng_boundary_bwd(const int32_t* __restrict__ slot_a, const float* __restrict__ msdf,
                                                               const float* __restrict__ tng, const float* __restrict__ g_tng,
                                                               int64_t n_boundary, int64_t n3, int64_t n_wt, float* __restrict__ g_t,
                                                               float* __restrict__ g_msdf) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= n_boundary) return;
  const Boundary e = boundary(slot_a, msdf, j, n3);
  if (!e.ok) return;                                               // both weights are the constant 0
  const V3 g = ld3(g_tng + (size_t)(n_wt + j) * 3);
  float* oa = g_t + (size_t)e.a * 3;
  float* ob = g_t + (size_t)e.b * 3;
  atomicAdd(oa, g.x * e.w0); atomicAdd(oa + 1, g.y * e.w0); atomicAdd(oa + 2, g.z * e.w0);
  atomicAdd(ob, g.x * e.w1); atomicAdd(ob + 1, g.y * e.w1); atomicAdd(ob + 2, g.z * e.w1);
  // w0 = -mb / D, w1 = ma / D, D = ma - mb:  d/dma = (<g,ta> - <g,tb>) mb / D^2,  d/dmb = (<g,tb> - <g,ta>) ma / D^2
  const float diff = dot(g, ld3(tng + (size_t)e.a * 3)) - dot(g, ld3(tng + (size_t)e.b * 3));
  const float D = e.ma - e.mb, inv2 = 1.f / (D * D);
  atomicAdd(g_msdf + e.a, diff * e.mb * inv2);
  atomicAdd(g_msdf + e.b, -diff * e.ma * inv2);
}
}  // namespace

extern "C" {

int gsb_tangents_fwd(const float* verts, const int32_t* faces, const float* normals, const float* msdf_wt, const int32_t* slot_a,
                     const float* atlas_lin, int64_t n_wt, int64_t n_faces, int64_t n_tri_polys, int64_t n_boundary,
                     int32_t atlas_n, float atlas_pad, float* acc, float* tng_aug, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_wt <= 0) return 0;
  if (atlas_n < 1 || n_faces < 0 || n_boundary < 0 || 3 * n_tri_polys > n_boundary || (n_boundary - 3 * n_tri_polys) % 4 != 0)
    return (int)cudaErrorInvalidValue;
  const Atlas atlas{atlas_lin, atlas_n, atlas_pad};
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float) * 4 * (size_t)n_wt, stream);
  if (e != cudaSuccess) return (int)e;
  if (n_faces > 0) k_tng_face<<<nblk(3 * n_faces), kThreads, 0, stream>>>(verts, faces, n_faces, atlas, acc);
  k_tng_vertex<<<nblk(n_wt), kThreads, 0, stream>>>(acc, normals, n_wt, tng_aug);
  if (n_boundary > 0)
    k_tng_boundary<<<nblk(n_boundary), kThreads, 0, stream>>>(slot_a, msdf_wt, n_boundary, 3 * n_tri_polys, n_wt, tng_aug);
  return (int)cudaGetLastError();
}

int gsb_tangents_bwd(const int32_t* faces, const float* normals, const float* msdf_wt, const int32_t* slot_a, const float* atlas_lin,
                     int64_t n_wt, int64_t n_faces, int64_t n_tri_polys, int64_t n_boundary, int32_t atlas_n, float atlas_pad,
                     const float* acc, const float* tng_aug, const float* g_tng_aug, float* g_t, float* g_verts, float* g_normals,
                     float* g_msdf, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_wt <= 0) return 0;
  if (atlas_n < 1 || n_faces < 0 || n_boundary < 0 || 3 * n_tri_polys > n_boundary || (n_boundary - 3 * n_tri_polys) % 4 != 0)
    return (int)cudaErrorInvalidValue;
  const Atlas atlas{atlas_lin, atlas_n, atlas_pad};
  cudaError_t e = cudaMemsetAsync(g_verts, 0, sizeof(float) * 3 * (size_t)n_wt, stream);
  if (e != cudaSuccess) return (int)e;
  k_tng_grad_init<<<nblk(n_wt), kThreads, 0, stream>>>(g_tng_aug, n_wt, g_t, g_msdf);
  if (n_boundary > 0)
    k_tng_boundary_bwd<<<nblk(n_boundary), kThreads, 0, stream>>>(slot_a, msdf_wt, tng_aug, g_tng_aug, n_boundary, 3 * n_tri_polys,
                                                                  n_wt, g_t, g_msdf);
  k_tng_vertex_bwd<<<nblk(n_wt), kThreads, 0, stream>>>(acc, normals, n_wt, g_t, g_normals);
  if (n_faces > 0) k_tng_face_bwd<<<nblk(n_faces), kThreads, 0, stream>>>(faces, n_faces, atlas, g_t, g_verts);
  return (int)cudaGetLastError();
}

}  // extern "C"
