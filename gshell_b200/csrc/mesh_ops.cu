// Smooth vertex normals of a triangle mesh, forward and adjoint, for sm_100a.
// Replaces mesh.auto_normals of the reference (render/mesh.py:212-237: cross products, three scatter_add_
// passes, normalisation, ~12 ATen kernels + their autograd twins) with 2 + 2 kernels.
// HBM/atomic bound: 36 B gathered + 9 red.global.add per face.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"
#include "vec.cuh"

using namespace gsb;

namespace {
constexpr int kThreads = 256;
inline int nblk(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

__global__ void __launch_bounds__(kThreads) k_face_splat(const float* __restrict__ v, const int32_t* __restrict__ tris,
                                                         int64_t n_faces, float* __restrict__ acc) {
  int64_t f = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (f >= n_faces) return;
  const int i0 = __ldg(tris + f * 3), i1 = __ldg(tris + f * 3 + 1), i2 = __ldg(tris + f * 3 + 2);
  const V3 p0 = ld3(v + (size_t)i0 * 3), p1 = ld3(v + (size_t)i1 * 3), p2 = ld3(v + (size_t)i2 * 3);
  const V3 n = cross(p1 - p0, p2 - p0);
  const int idx[3] = {i0, i1, i2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float* a = acc + (size_t)idx[k] * 3;
    atomicAdd(a, n.x); atomicAdd(a + 1, n.y); atomicAdd(a + 2, n.z);
  }
}

// n = acc / sqrt(max(|acc|^2, 1e-20)), degenerate (|acc|^2 <= 1e-20) -> (0,0,1)   (mesh.py:231-232)
__global__ void __launch_bounds__(kThreads) k_normalize(const float* __restrict__ acc, int64_t n_verts, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_verts) return;
  V3 a = ld3(acc + i * 3);
  float d = dot(a, a);
  if (!(d > 1e-20f)) { a = v3(0.f, 0.f, 1.f); d = 1.f; }
  st3(out + i * 3, a / sqrtf(fmaxf(d, 1e-20f)));
}

__global__ void __launch_bounds__(kThreads) k_normalize_bwd(const float* __restrict__ acc, const float* __restrict__ g_n,
                                                            int64_t n_verts, float* __restrict__ g_acc) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_verts) return;
  const V3 a = ld3(acc + i * 3), g = ld3(g_n + i * 3);
  const float d = dot(a, a);
  V3 r = v3(0.f);
  if (d > 1e-20f) {
    const float inv = rsqrtf(d);
    r = (g - a * (dot(a, g) / d)) * inv;
  }
  st3(g_acc + i * 3, r);
}

__global__ void __launch_bounds__(kThreads) k_face_splat_bwd(const float* __restrict__ v, const int32_t* __restrict__ tris,
                                                             const float* __restrict__ g_acc, int64_t n_faces,
                                                             float* __restrict__ g_v) {
  int64_t f = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (f >= n_faces) return;
  const int i0 = __ldg(tris + f * 3), i1 = __ldg(tris + f * 3 + 1), i2 = __ldg(tris + f * 3 + 2);
  const V3 p0 = ld3(v + (size_t)i0 * 3), p1 = ld3(v + (size_t)i1 * 3), p2 = ld3(v + (size_t)i2 * 3);
  const V3 g = ld3(g_acc + (size_t)i0 * 3) + ld3(g_acc + (size_t)i1 * 3) + ld3(g_acc + (size_t)i2 * 3);
  V3 ga = v3(0.f), gb = v3(0.f);
  cross_bwd(p1 - p0, p2 - p0, g, ga, gb);
  const V3 g0 = -(ga + gb);
  float* o0 = g_v + (size_t)i0 * 3;
  float* o1 = g_v + (size_t)i1 * 3;
  float* o2 = g_v + (size_t)i2 * 3;
  atomicAdd(o0, g0.x); atomicAdd(o0 + 1, g0.y); atomicAdd(o0 + 2, g0.z);
  atomicAdd(o1, ga.x); atomicAdd(o1 + 1, ga.y); atomicAdd(o1 + 2, ga.z);
  atomicAdd(o2, gb.x); atomicAdd(o2 + 1, gb.y); atomicAdd(o2 + 2, gb.z);
}
}  // namespace

extern "C" {

int gsb_vertex_normals_fwd(const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_faces, float* acc,
                           float* normals, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_verts == 0) return 0;
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float) * 3 * (size_t)n_verts, stream);
  if (e != cudaSuccess) return (int)e;
  if (n_faces > 0) k_face_splat<<<nblk(n_faces), kThreads, 0, stream>>>(verts, tris, n_faces, acc);
  k_normalize<<<nblk(n_verts), kThreads, 0, stream>>>(acc, n_verts, normals);
  return (int)cudaGetLastError();
}

int gsb_vertex_normals_bwd(const float* verts, const int32_t* tris, const float* acc, const float* g_normals,
                           int64_t n_verts, int64_t n_faces, float* g_acc, float* g_verts, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_verts == 0) return 0;
  cudaError_t e = cudaMemsetAsync(g_verts, 0, sizeof(float) * 3 * (size_t)n_verts, stream);
  if (e != cudaSuccess) return (int)e;
  k_normalize_bwd<<<nblk(n_verts), kThreads, 0, stream>>>(acc, g_normals, n_verts, g_acc);
  if (n_faces > 0) k_face_splat_bwd<<<nblk(n_faces), kThreads, 0, stream>>>(verts, tris, g_acc, n_faces, g_verts);
  return (int)cudaGetLastError();
}

}  // extern "C"
