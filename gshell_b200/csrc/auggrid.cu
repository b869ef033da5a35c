// Decode of a generated "augmented grid" into an open mesh, for sm_100a (no gradients, as in the reference).
//
// Replaces GShell_Tets.marching_from_auggrid (reference geometry/gshell_tets.py:446-629): there ~90 ATen ops with a row-wise
// unique over the valid tets' edges, ~20 boolean-mask compactions (each a host sync) and six gathers through look-up tables; here
// five kernels around two prefix sums (torch.cumsum on int32 flags, done by the host layer between the calls):
//
//   gsb_auggrid_edge_flags   one flag per edge of the static sorted edge table: does the SDF change sign on it?
//        -> scan: vertex id of a crossing edge = its rank among the crossing edges (the order the reference's `unique` yields)
//   gsb_auggrid_vertices     per crossing edge: position from the generated interpolation coefficient, canonical mid-point, mSDF sign
//   gsb_auggrid_classify     per tet: SDF case, 1- or 2-triangle polygon, mSDF cut code -> eight 0/1 rows (two polygon classes,
//                            six cut groups)
//        -> scan of the eight rows: output rows of every tet
//   gsb_auggrid_emit         per tet: watertight faces, tet id, boundary vertices (occupancy-grid weights), cut faces in the
//                            reference's six groups
//   gsb_auggrid_boundary_attr  per boundary vertex: any per-vertex attribute (the tangents) with the same weights
//
// One independent thread per element in every kernel (also compiled as host code by the CPU tests, tests/native/host_kernels.py).
// -fmad=false: positions reproduce the separately rounded products of the PyTorch ops they replace.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"
#include "vec.cuh"

using namespace gsb;

namespace {
constexpr int kThreads = 256;
inline int nblk(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

struct Grid3 {                      // dense [nx, ny, nz] float array, C order
  const float* v;
  int nx, ny, nz;
};
__device__ __forceinline__ float grid_at(const Grid3& g, int x, int y, int z) {
  x = min(max(x, 0), g.nx - 1); y = min(max(y, 0), g.ny - 1); z = min(max(z, 0), g.nz - 1);   // out of range is an error in the reference
  return __ldg(g.v + ((size_t)x * g.ny + y) * g.nz + z);
}

struct Luts {                       // int32, as gshell_b200/geometry/mt_luts.py lists them (negative entries clamped to 0)
  const int32_t *tri, *loop, *ntri, *cut3, *cut4, *ncut3, *ncut4;     // [16,6] [16,4] [16] [8,6] [16,12] [8] [16]
};

__device__ __forceinline__ bool inside(const float* __restrict__ sdf, int v) { return __ldg(sdf + v) > 0.f; }

__global__ void __launch_bounds__(kThreads) k_aug_edge_flags(const float* __restrict__ sdf, const int32_t* __restrict__ edge_v,
                                                             int64_t n_edges, int32_t* __restrict__ flags) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= n_edges) return;
  flags[e] = inside(sdf, __ldg(edge_v + e * 2)) != inside(sdf, __ldg(edge_v + e * 2 + 1)) ? 1 : 0;
}

// verts = p_hi * c + p_lo * (1 - c), c = clamp(coeff[mid], 0, 1), mid = integer mid-point of the canonical edge (:474-491)
__global__ void __launch_bounds__(kThreads) k_aug_vertices(const float* __restrict__ pos, const float* __restrict__ disc,
                                                           const int32_t* __restrict__ edge_v, const int32_t* __restrict__ flags,
                                                           const int32_t* __restrict__ incl, int64_t n_edges, Grid3 coeff, Grid3 msign,
                                                           float* __restrict__ verts, float* __restrict__ cano, float* __restrict__ m_vert) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= n_edges || !__ldg(flags + e)) return;
  const int64_t v = __ldg(incl + e) - 1;
  const int lo = __ldg(edge_v + e * 2), hi = __ldg(edge_v + e * 2 + 1);
  const V3 c0 = ld3(disc + (size_t)lo * 3), c1 = ld3(disc + (size_t)hi * 3);
  const V3 mid = (c0 + c1) / 2.0f;
  const int mx = (int)mid.x, my = (int)mid.y, mz = (int)mid.z;
  const float c = clampf(grid_at(coeff, mx, my, mz), 0.f, 1.f);
  st3(verts + v * 3, ld3(pos + (size_t)hi * 3) * c + ld3(pos + (size_t)lo * 3) * (1.f - c));
  st3(cano + v * 3, mid);
  m_vert[v] = grid_at(msign, mx, my, mz);
}

struct TetCase { int code, ntri; int vmap[6]; };
__device__ __forceinline__ TetCase tet_case(const float* __restrict__ sdf, const int32_t* __restrict__ tet_v,
                                            const int32_t* __restrict__ tet_e, const int32_t* __restrict__ flags,
                                            const int32_t* __restrict__ incl, const Luts& L, int64_t t) {
  TetCase r;
  r.code = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.code |= inside(sdf, __ldg(tet_v + t * 4 + k)) ? (1 << k) : 0;
  r.ntri = __ldg(L.ntri + r.code);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int e = __ldg(tet_e + t * 6 + k);
    r.vmap[k] = __ldg(flags + e) ? __ldg(incl + e) - 1 : -1;
  }
  return r;
}
// first vertices of the polygon's edges (the loop closes on entry 0) and the mSDF cut code of the polygon (:520-523)
__device__ __forceinline__ int loop_code(const TetCase& c, const Luts& L, const float* __restrict__ m_vert, int* a) {
  const int n = c.ntri == 1 ? 3 : 4;
  int code = 0;
  for (int i = 0; i < n; ++i) {
    a[i] = c.vmap[__ldg(L.loop + c.code * 4 + i)];
    code = (code << 1) | (__ldg(m_vert + a[i]) > 0.f ? 1 : 0);
  }
  return code;
}

// rows[0] / rows[1]: the tet makes a triangle / quad polygon; rows[2 + g]: its polygon falls into cut group g
// (triangle -> 1, 2 faces; quad -> 1, 2, 3, 4 faces).  Every row entry of every tet is written.
__global__ void __launch_bounds__(kThreads) k_aug_classify(const float* __restrict__ sdf, const int32_t* __restrict__ tet_v,
                                                           const int32_t* __restrict__ tet_e, const int32_t* __restrict__ flags,
                                                           const int32_t* __restrict__ incl, const float* __restrict__ m_vert,
                                                           int64_t n_tets, Luts L, int32_t* __restrict__ rows) {
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_tets) return;
  const TetCase c = tet_case(sdf, tet_v, tet_e, flags, incl, L, t);
  int group = -1;
  if (c.ntri > 0) {
    int a[4];
    const int code = loop_code(c, L, m_vert, a);
    const int k = c.ntri == 1 ? __ldg(L.ncut3 + code) : __ldg(L.ncut4 + code);
    if (k > 0) group = c.ntri == 1 ? k - 1 : 1 + k;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int on = r == 0 ? c.ntri == 1 : (r == 1 ? c.ntri == 2 : group == r - 2);
    rows[(size_t)r * n_tets + t] = on ? 1 : 0;
  }
}

struct EmitSizes {
  int64_t n_wt, n_one, n_two;
  int64_t group_start[6];            // first row of each cut group in faces_aug
};

__global__ void __launch_bounds__(kThreads) k_aug_emit(const float* __restrict__ sdf, const int32_t* __restrict__ tet_v,
                                                       const int32_t* __restrict__ tet_e, const int32_t* __restrict__ flags,
                                                       const int32_t* __restrict__ incl, const float* __restrict__ m_vert,
                                                       const float* __restrict__ verts, const float* __restrict__ cano,
                                                       const int32_t* __restrict__ rows_incl, int64_t n_tets, Luts L, Grid3 occ, EmitSizes z,
                                                       int32_t* __restrict__ faces_wt, int32_t* __restrict__ tet_ids,
                                                       float* __restrict__ b_pos, int32_t* __restrict__ b_ab, float* __restrict__ b_w,
                                                       int32_t* __restrict__ faces_aug) {
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_tets) return;
  const TetCase c = tet_case(sdf, tet_v, tet_e, flags, incl, L, t);
  if (c.ntri == 0) return;
  const bool quad = c.ntri == 2;
  const int n = quad ? 4 : 3;
  const int64_t rank = __ldg(rows_incl + (size_t)(quad ? 1 : 0) * n_tets + t) - 1;      // among the tets of its polygon class
  // watertight faces and the tet id, triangles of all 1-triangle tets first (:499-510)
  const int64_t frow = quad ? z.n_one + 2 * rank : rank;
  for (int i = 0; i < 3 * c.ntri; ++i) faces_wt[frow * 3 + i] = c.vmap[__ldg(L.tri + c.code * 6 + i)];
  tet_ids[quad ? z.n_one + rank : rank] = (int32_t)t;
  // boundary vertices: one per polygon edge (a_i, a_{i+1}), weights from the generated occupancy grid at the edge's mid-point in
  // the doubled grid; the weight order follows the lexicographic order of the canonical end points (:536-575)
  int a[4];
  const int code = loop_code(c, L, m_vert, a);
  const int64_t slot0 = quad ? 3 * z.n_one + 4 * rank : 3 * rank;
  for (int i = 0; i < n; ++i) {
    const int va = a[i], vb = a[(i + 1) % n];
    const V3 ca = ld3(cano + (size_t)va * 3), cb = ld3(cano + (size_t)vb * 3);
    const V3 loc = ((ca + cb) / 2.0f) * 2.0f;
    const float co = grid_at(occ, (int)loc.x, (int)loc.y, (int)loc.z) * 0.5f + 0.5f;
    const V3 d = ca - cb;
    const float key = 16.f * (float)((d.x > 0.f) - (d.x < 0.f)) + 4.f * (float)((d.y > 0.f) - (d.y < 0.f)) + (float)((d.z > 0.f) - (d.z < 0.f));
    const float w0 = key >= 0.f ? co : 1.f - co, w1 = key >= 0.f ? 1.f - co : co;
    const int64_t s = slot0 + i;
    st3(b_pos + s * 3, ld3(verts + (size_t)va * 3) * w0 + ld3(verts + (size_t)vb * 3) * w1);
    b_ab[s * 2] = va; b_ab[s * 2 + 1] = vb;
    b_w[s * 2] = w0; b_w[s * 2 + 1] = w1;
  }
  // cut faces (:596-620): ids = the polygon's watertight vertices followed by its boundary vertices
  const int k = quad ? __ldg(L.ncut4 + code) : __ldg(L.ncut3 + code);
  if (k == 0) return;
  const int g = quad ? 1 + k : k - 1;
  const int64_t grank = __ldg(rows_incl + (size_t)(2 + g) * n_tets + t) - 1;
  int32_t* out = faces_aug + (z.group_start[g] + grank * k) * 3;
  const int32_t* table = quad ? L.cut4 + code * 12 : L.cut3 + code * 6;
  for (int i = 0; i < 3 * k; ++i) {
    const int id = __ldg(table + i);
    out[i] = id < n ? a[id] : (int32_t)(z.n_wt + slot0 + (id - n));
  }
}

__global__ void __launch_bounds__(kThreads) k_aug_boundary_attr(const float* __restrict__ attr, const int32_t* __restrict__ b_ab,
                                                                const float* __restrict__ b_w, int64_t n_boundary, float* __restrict__ out) {
  const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (s >= n_boundary) return;
  st3(out + s * 3, ld3(attr + (size_t)__ldg(b_ab + s * 2) * 3) * __ldg(b_w + s * 2) + ld3(attr + (size_t)__ldg(b_ab + s * 2 + 1) * 3) * __ldg(b_w + s * 2 + 1));
}

inline Luts luts_of(const int32_t* const* p) { return Luts{p[0], p[1], p[2], p[3], p[4], p[5], p[6]}; }
}  // namespace

extern "C" {

int gsb_auggrid_edge_flags(const float* sdf, const int32_t* edge_v, int64_t n_edges, int32_t* flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_edges <= 0) return 0;
  k_aug_edge_flags<<<nblk(n_edges), kThreads, 0, stream>>>(sdf, edge_v, n_edges, flags);
  return (int)cudaGetLastError();
}

int gsb_auggrid_vertices(const float* pos, const float* verts_discretized, const int32_t* edge_v, const int32_t* flags,
                         const int32_t* flags_incl, int64_t n_edges, const float* coeff_grid, const float* msdf_sign_grid,
                         int32_t gx, int32_t gy, int32_t gz, float* verts, float* verts_cano, float* msdf_vert, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_edges <= 0) return 0;
  if (gx < 1 || gy < 1 || gz < 1) return (int)cudaErrorInvalidValue;
  const Grid3 coeff{coeff_grid, gx, gy, gz}, msign{msdf_sign_grid, gx, gy, gz};
  k_aug_vertices<<<nblk(n_edges), kThreads, 0, stream>>>(pos, verts_discretized, edge_v, flags, flags_incl, n_edges, coeff, msign, verts,
                                                         verts_cano, msdf_vert);
  return (int)cudaGetLastError();
}

int gsb_auggrid_classify(const float* sdf, const int32_t* tet_v, const int32_t* tet_e, const int32_t* flags, const int32_t* flags_incl,
                         const float* msdf_vert, int64_t n_tets, const int32_t* const* luts7_host, int32_t* rows, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_tets <= 0) return 0;
  k_aug_classify<<<nblk(n_tets), kThreads, 0, stream>>>(sdf, tet_v, tet_e, flags, flags_incl, msdf_vert, n_tets, luts_of(luts7_host), rows);
  return (int)cudaGetLastError();
}

int gsb_auggrid_emit(const float* sdf, const int32_t* tet_v, const int32_t* tet_e, const int32_t* flags, const int32_t* flags_incl,
                     const float* msdf_vert, const float* verts, const float* verts_cano, const int32_t* rows_incl, int64_t n_tets,
                     const int32_t* const* luts7_host, const float* occgrid, int32_t ox, int32_t oy, int32_t oz, int64_t n_wt,
                     const int64_t* totals8_host, int32_t* faces_wt, int32_t* tet_ids, float* boundary_pos, int32_t* boundary_ab, float* boundary_w,
                     int32_t* faces_aug, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_tets <= 0) return 0;
  if (ox < 1 || oy < 1 || oz < 1) return (int)cudaErrorInvalidValue;
  EmitSizes z;
  z.n_wt = n_wt; z.n_one = totals8_host[0]; z.n_two = totals8_host[1];
  static const int faces_per_poly[6] = {1, 2, 1, 2, 3, 4};
  int64_t row = 0;
  for (int g = 0; g < 6; ++g) { z.group_start[g] = row; row += totals8_host[2 + g] * faces_per_poly[g]; }
  const Grid3 occ{occgrid, ox, oy, oz};
  k_aug_emit<<<nblk(n_tets), kThreads, 0, stream>>>(sdf, tet_v, tet_e, flags, flags_incl, msdf_vert, verts, verts_cano, rows_incl, n_tets,
                                                    luts_of(luts7_host), occ, z, faces_wt, tet_ids, boundary_pos, boundary_ab, boundary_w,
                                                    faces_aug);
  return (int)cudaGetLastError();
}

int gsb_auggrid_boundary_attr(const float* attr, const int32_t* boundary_ab, const float* boundary_w, int64_t n_boundary, float* out,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_boundary <= 0) return 0;
  k_aug_boundary_attr<<<nblk(n_boundary), kThreads, 0, stream>>>(attr, boundary_ab, boundary_w, n_boundary, out);
  return (int)cudaGetLastError();
}

}  // extern "C"
