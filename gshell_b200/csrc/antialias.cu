// Silhouette antialiasing with gradients to the vertex positions -- the operator the reference gets from nvdiffrast
// (`dr.antialias`, reference render/render.py:352-359, twelve calls per iteration).  It is the ONLY path by which the coverage
// (alpha) of the rendered image depends differentiably on geometry: F.mse_loss(shaded.alpha, target.alpha) in tick() moves the
// surface through this operator.  nvdiffrast is a third-party, un-vendored dependency (reference README.md:38) that is absent
// here, so this is an own implementation of its documented algorithm -- PARITY UNPINNED, see DESIGN.md:
//
//   for every horizontally / vertically adjacent pixel pair whose triangle ids differ
//     front  = the triangle of the pixel nearer to the camera (background counts as infinitely far)
//     follow the segment between the two pixel centres from `front` through its neighbours (<= 4 hops) to the first edge that
//     is a silhouette -- it has one adjacent triangle, or its two adjacent triangles face opposite ways on screen --
//     crossed at parameter t in [0,1] from the front pixel's centre; an edge only serves the pairs along its dominant normal
//     direction (|dy| >= |dx| edges blend horizontal pairs, the others vertical pairs), so that no crossing is counted twice
//     t > 0.5: the front surface covers (t - 0.5) of the far pixel  ->  far  += (t - 0.5) (front - far)
//     t < 0.5: the far surface shows in (0.5 - t) of the front pixel ->  front += (0.5 - t) (far - front)
//   backward: d/d colours through the blend weights, d/dt -> the edge's two vertices in clip space (x, y, w).
//
// The pair analysis (edge hash of the mesh, silhouette test, crossing) runs ONCE per frame and leaves a compact list of work
// items that every antialiased buffer, forward and backward, replays: the mesh changes every iteration, so the edge ->
// opposite-vertex hash is rebuilt per iteration too (3 F inserts; count -> no scan needed, open addressing).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"

namespace {
constexpr int kThreads = 256;
inline int nblk(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

struct EdgeSlot {
  unsigned long long key;   // (lo << 32 | hi) + 1; 0 = empty
  int op0, op1;             // opposite vertices of the (up to two) triangles on this edge; -1 = none
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ unsigned long long edge_key(int a, int b) {
  const unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
  return (((unsigned long long)lo << 32) | hi) + 1ull;
}

__global__ void __launch_bounds__(kThreads) k_hash_clear(EdgeSlot* __restrict__ tab, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < n) { tab[i].key = 0ull; tab[i].op0 = -1; tab[i].op1 = -1; }
}

__global__ void __launch_bounds__(kThreads) k_hash_build(const int32_t* __restrict__ tris, int64_t F, EdgeSlot* __restrict__ tab,
                                                         uint64_t mask) {
  const int64_t f = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (f >= F) return;
  const int v[3] = {__ldg(tris + f * 3), __ldg(tris + f * 3 + 1), __ldg(tris + f * 3 + 2)};
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const int a = v[e], b = v[(e + 1) % 3], op = v[(e + 2) % 3];
    if (a == b) continue;
    const unsigned long long key = edge_key(a, b);
    uint64_t h = mix64(key) & mask;
    for (;;) {
      const unsigned long long prev = atomicCAS(&tab[h].key, 0ull, key);
      if (prev == 0ull || prev == key) {
        if (atomicCAS(&tab[h].op0, -1, op) != -1) atomicCAS(&tab[h].op1, -1, op);     // a third triangle on the edge is ignored
        break;
      }
      h = (h + 1) & mask;
    }
  }
}

__device__ __forceinline__ bool hash_find(const EdgeSlot* __restrict__ tab, uint64_t mask, int a, int b, int& op0, int& op1) {
  const unsigned long long key = edge_key(a, b);
  uint64_t h = mix64(key) & mask;
  for (int probe = 0; probe < 1 << 20; ++probe) {
    const unsigned long long k = tab[h].key;
    if (k == key) { op0 = tab[h].op0; op1 = tab[h].op1; return true; }
    if (k == 0ull) return false;
    h = (h + 1) & mask;
  }
  return false;
}

// screen position in pixel units (pixel (i, j) has its centre at (i + 0.5, j + 0.5)) and 1/w
struct Scr { float x, y, iw; };
__device__ __forceinline__ Scr to_screen(float4 c, float W, float H) {
  Scr s;
  s.iw = 1.f / c.w;
  s.x = (c.x * s.iw * 0.5f + 0.5f) * W;
  s.y = (c.y * s.iw * 0.5f + 0.5f) * H;
  return s;
}
__device__ __forceinline__ float area2(Scr a, Scr b, Scr c) { return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x); }

#ifndef GSB_AA_HOPS
#define GSB_AA_HOPS 4
#endif
constexpr int kHops = GSB_AA_HOPS;   // neighbours the pair analysis may walk through before it gives up (0 = rasterised triangle only)

struct Item {        // 32 bytes
  int pix_front, pix_far;   // flat pixel indices (batch included)
  int va, vb;               // the silhouette edge (vertex ids)
  float t;                  // crossing parameter from the front pixel's centre
  int axis;                 // 0: the pair is horizontal (the edge crossing moves along x), 1: vertical
  float dir;                // +1 if far = front + 1 along the axis, -1 otherwise
  int batch;
};

// one thread per pixel: the pair with its right neighbour and the pair with the neighbour below
__global__ void __launch_bounds__(kThreads) k_analyse(const float4* __restrict__ rast, const float4* __restrict__ clip, const int32_t* __restrict__ tris,
                                                      const EdgeSlot* __restrict__ tab, uint64_t mask, int n_verts, int W, int H, int64_t n_pix,
                                                      Item* __restrict__ items, int* __restrict__ n_items, int cap) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_pix) return;
  const int px = (int)(i % W), py = (int)((i / W) % H), b = (int)(i / ((int64_t)W * H));
  const float4 r0 = __ldg(rast + i);
  const int id0 = (int)r0.w;
#pragma unroll
  for (int axis = 0; axis < 2; ++axis) {
    if (axis == 0 ? px + 1 >= W : py + 1 >= H) continue;
    const int64_t j = axis == 0 ? i + 1 : i + W;
    const float4 r1 = __ldg(rast + j);
    const int id1 = (int)r1.w;
    if (id0 == id1) continue;
    // nearer pixel = front (NDC depth grows with distance); background is behind everything
    const bool first_front = id1 == 0 || (id0 != 0 && r0.z <= r1.z);
    const int tri = (first_front ? id0 : id1) - 1;
    const int64_t pf = first_front ? i : j, pb = first_front ? j : i;
    const float dir = first_front ? 1.f : -1.f;
    const float fx = (float)(first_front ? px : px + (axis == 0)) + 0.5f, fy = (float)(first_front ? py : py + (axis == 1)) + 0.5f;
    const size_t off = (size_t)b * n_verts;
    const int v[3] = {__ldg(tris + (size_t)tri * 3), __ldg(tris + (size_t)tri * 3 + 1), __ldg(tris + (size_t)tri * 3 + 2)};
    Scr s[3];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 c = __ldg(clip + off + v[k]);
      ok &= c.w > 1e-8f;
      s[k] = to_screen(c, (float)W, (float)H);
    }
    if (!ok) continue;
    // Walk from the front pixel's triangle towards the far pixel across non-silhouette edges (at most kHops neighbours): next
    // to a contour the triangle that owns the silhouette edge is foreshortened to a fraction of a pixel and rarely covers a
    // pixel centre itself, so looking only at the rasterised triangle (as nvdiffrast does) misses most contour crossings of a
    // finely tessellated smooth surface.
    int ta = v[0], tb = v[1], tc = v[2];             // current triangle, mesh orientation
    Scr sa = s[0], sb = s[1], sc = s[2];
    int entry = -1;                                  // edge (0: a-b, 1: b-c, 2: c-a) the walk came in through
    float t_prev = -1.f, best_t = -1.f;
    int best_a = 0, best_b = 0;
    for (int hop = 0; hop <= kHops; ++hop) {
      const float a_self = area2(sa, sb, sc);
      const int ev[3][2] = {{ta, tb}, {tb, tc}, {tc, ta}};
      const Scr es[3][2] = {{sa, sb}, {sb, sc}, {sc, sa}};
      const int opp[3] = {tc, ta, tb};
      int hit = -1;
      float t_hit = 2.f;
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        if (e == entry) continue;
        const Scr P = es[e][0], Q = es[e][1];
        // crossing of the edge with the segment between the two pixel centres
        const float pc = axis == 0 ? P.y - fy : P.x - fx, qc = axis == 0 ? Q.y - fy : Q.x - fx;      // across the segment's line
        if ((pc > 0.f) == (qc > 0.f)) continue;
        const float sp = pc / (pc - qc);
        const float along = axis == 0 ? (P.x + sp * (Q.x - P.x)) - fx : (P.y + sp * (Q.y - P.y)) - fy;
        const float t = along * dir;
        if (!(t >= 0.f && t <= 1.f) || t < t_prev) continue;
        if (t < t_hit) { t_hit = t; hit = e; }       // the first edge the segment leaves the triangle through
      }
      if (hit < 0) break;
      const Scr P = es[hit][0], Q = es[hit][1];
      int o0, o1;
      if (!hash_find(tab, mask, ev[hit][0], ev[hit][1], o0, o1)) break;
      int other = -1;
      Scr so = P;
      bool sil = o1 < 0;                             // open boundary edge
      if (!sil) {
        other = o0 == opp[hit] ? o1 : o0;
        const float4 co = __ldg(clip + off + other);
        if (!(co.w > 1e-8f)) break;
        so = to_screen(co, (float)W, (float)H);
        // the neighbour across the edge, (Q, P, other), keeps the mesh orientation; opposite screen orientation = fold
        sil = (a_self > 0.f) != (area2(Q, P, so) > 0.f);
      }
      if (sil) {
        // each edge blends only the pixel pairs along its dominant normal direction: a diagonal edge crossing both the
        // horizontal and the vertical pair of a pixel would otherwise be counted twice
        const float ex = fabsf(Q.x - P.x), ey = fabsf(Q.y - P.y);
        if (axis == 0 ? ey >= ex : ex > ey) { best_t = t_hit; best_a = ev[hit][0]; best_b = ev[hit][1]; }
        break;
      }
      // hop into the neighbour (Q, P, other); it was entered through its edge 0 (Q - P)
      ta = ev[hit][1]; tb = ev[hit][0]; tc = other;
      sa = Q; sb = P; sc = so;
      entry = 0;
      t_prev = t_hit;
    }
    if (best_t < 0.f) continue;
    const int slot = atomicAdd(n_items, 1);
    if (slot < cap) {
      Item it;
      it.pix_front = (int)pf; it.pix_far = (int)pb; it.va = best_a; it.vb = best_b; it.t = best_t; it.axis = axis; it.dir = dir; it.batch = b;
      items[slot] = it;
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_aa_fwd(const float* __restrict__ color, const Item* __restrict__ items, const int* __restrict__ n_items,
                                                     int cap, int C, float* __restrict__ out) {
  const int n = min(*n_items, cap);
  for (int k = blockIdx.x * kThreads + threadIdx.x; k < n; k += gridDim.x * kThreads) {
    const Item it = items[k];
    const float t = it.t;
    const int dst = t > 0.5f ? it.pix_far : it.pix_front, src = t > 0.5f ? it.pix_front : it.pix_far;
    const float w = fabsf(t - 0.5f);
    for (int c = 0; c < C; ++c) {
      const float d = __ldg(color + (size_t)src * C + c) - __ldg(color + (size_t)dst * C + c);
      if (d != 0.f) atomicAdd(out + (size_t)dst * C + c, w * d);
    }
  }
}

// g_color is pre-set to g_out (the pass-through part); g_clip is accumulated
__global__ void __launch_bounds__(kThreads) k_aa_bwd(const float* __restrict__ color, const float* __restrict__ g_out, const Item* __restrict__ items,
                                                     const int* __restrict__ n_items, int cap, int C, const float4* __restrict__ clip, int n_verts,
                                                     int W, int H, float* __restrict__ g_color, float* __restrict__ g_clip) {
  const int n = min(*n_items, cap);
  for (int k = blockIdx.x * kThreads + threadIdx.x; k < n; k += gridDim.x * kThreads) {
    const Item it = items[k];
    const float t = it.t;
    const bool to_far = t > 0.5f;
    const int dst = to_far ? it.pix_far : it.pix_front, src = to_far ? it.pix_front : it.pix_far;
    const float w = fabsf(t - 0.5f);
    float g_w = 0.f;
    for (int c = 0; c < C; ++c) {
      const float g = __ldg(g_out + (size_t)dst * C + c);
      if (g == 0.f) continue;
      const float d = __ldg(color + (size_t)src * C + c) - __ldg(color + (size_t)dst * C + c);
      g_w += g * d;
      if (g_color) {
        atomicAdd(g_color + (size_t)src * C + c, w * g);
        atomicAdd(g_color + (size_t)dst * C + c, -w * g);
      }
    }
    if (!g_clip || g_w == 0.f) continue;
    const float g_t = to_far ? g_w : -g_w;                       // w = |t - 0.5|
    // t = dir * (x*(P, Q) - f) with x* the crossing along the pair's axis
    const size_t off = (size_t)it.batch * n_verts;
    const float4 cP = __ldg(clip + off + it.va), cQ = __ldg(clip + off + it.vb);
    const Scr P = to_screen(cP, (float)W, (float)H), Q = to_screen(cQ, (float)W, (float)H);
    const int fpx = it.pix_front % W, fpy = (it.pix_front / W) % H;
    const float fx = (float)fpx + 0.5f, fy = (float)fpy + 0.5f;
    // u = coordinate along the pair's axis, v = across it
    const float Pu = it.axis == 0 ? P.x : P.y, Pv = it.axis == 0 ? P.y - fy : P.x - fx;
    const float Qu = it.axis == 0 ? Q.x : Q.y, Qv = it.axis == 0 ? Q.y - fy : Q.x - fx;
    const float den = Pv - Qv;
    if (fabsf(den) < 1e-12f) continue;
    const float sp = Pv / den;
    const float g_along = g_t * it.dir;
    // along = Pu + sp (Qu - Pu);  sp = Pv / (Pv - Qv)
    const float g_Pu = g_along * (1.f - sp), g_Qu = g_along * sp;
    const float g_sp = g_along * (Qu - Pu);
    const float g_Pv = g_sp * (-Qv) / (den * den), g_Qv = g_sp * Pv / (den * den);
    const float gPx = it.axis == 0 ? g_Pu : g_Pv, gPy = it.axis == 0 ? g_Pv : g_Pu;
    const float gQx = it.axis == 0 ? g_Qu : g_Qv, gQy = it.axis == 0 ? g_Qv : g_Qu;
    // screen (pixel units) -> clip: X = (x / w * 0.5 + 0.5) W
    float* oP = g_clip + (off + it.va) * 4;
    float* oQ = g_clip + (off + it.vb) * 4;
    const float kx = 0.5f * (float)W, ky = 0.5f * (float)H;
    atomicAdd(oP, gPx * kx * P.iw); atomicAdd(oP + 1, gPy * ky * P.iw);
    atomicAdd(oP + 3, -(gPx * kx * cP.x + gPy * ky * cP.y) * P.iw * P.iw);
    atomicAdd(oQ, gQx * kx * Q.iw); atomicAdd(oQ + 1, gQy * ky * Q.iw);
    atomicAdd(oQ + 3, -(gQx * kx * cQ.x + gQy * ky * cQ.y) * Q.iw * Q.iw);
  }
}

}  // namespace

extern "C" {

/* slots of the edge hash for a mesh of n_faces triangles (power of two, load <= 0.5); 16 bytes per slot */
int64_t gsb_antialias_hash_slots(int64_t n_faces) {
  int64_t n = 16;
  while (n < 6 * n_faces) n <<= 1;
  return n;
}
size_t gsb_antialias_item_bytes(void) { return sizeof(Item); }

/* Pair analysis of one frame: builds the edge hash of `tris` in hash_ws (gsb_antialias_hash_slots(F) * 16 bytes) and appends one
 * work item per silhouette crossing to items (item_cap * gsb_antialias_item_bytes() bytes); *n_items (device int, zeroed here)
 * counts them -- items past item_cap are dropped and the count says so.  item_cap = 2 * B * H * W always suffices. */
int gsb_antialias_analyse(const float* rast, const float* clip, const int32_t* tris, int64_t n_batch, int64_t H, int64_t W, int64_t n_verts,
                          int64_t n_faces, void* hash_ws, void* items, int32_t* n_items, int64_t item_cap, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t slots = gsb_antialias_hash_slots(n_faces), n_pix = n_batch * H * W;
  cudaError_t e = cudaMemsetAsync(n_items, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return (int)e;
  if (n_pix == 0 || n_faces == 0) return 0;
  k_hash_clear<<<nblk(slots), kThreads, 0, stream>>>((EdgeSlot*)hash_ws, slots);
  k_hash_build<<<nblk(n_faces), kThreads, 0, stream>>>(tris, n_faces, (EdgeSlot*)hash_ws, (uint64_t)(slots - 1));
  k_analyse<<<nblk(n_pix), kThreads, 0, stream>>>((const float4*)rast, (const float4*)clip, tris, (const EdgeSlot*)hash_ws, (uint64_t)(slots - 1),
                                                 (int)n_verts, (int)W, (int)H, n_pix, (Item*)items, n_items,
                                                 (int)(item_cap < 0x7fffffff ? item_cap : 0x7fffffff));
  return (int)cudaGetLastError();
}

/* out must hold a copy of color ([n_pix, C]); the blends are added on top */
int gsb_antialias_fwd(const float* color, const void* items, const int32_t* n_items, int64_t item_cap, int64_t n_channels, float* out,
                      void* stream_) {
  k_aa_fwd<<<148 * 8, kThreads, 0, (cudaStream_t)stream_>>>(color, (const Item*)items, n_items, (int)(item_cap < 0x7fffffff ? item_cap : 0x7fffffff),
                                                           (int)n_channels, out);
  return (int)cudaGetLastError();
}

/* g_color (or NULL) must hold a copy of g_out; g_clip [B,V,4] (or NULL) is accumulated */
int gsb_antialias_bwd(const float* color, const float* g_out, const void* items, const int32_t* n_items, int64_t item_cap, int64_t n_channels,
                      const float* clip, int64_t n_verts, int64_t H, int64_t W, float* g_color, float* g_clip, void* stream_) {
  k_aa_bwd<<<148 * 8, kThreads, 0, (cudaStream_t)stream_>>>(color, g_out, (const Item*)items, n_items,
                                                           (int)(item_cap < 0x7fffffff ? item_cap : 0x7fffffff), (int)n_channels,
                                                           (const float4*)clip, (int)n_verts, (int)W, (int)H, g_color, g_clip);
  return (int)cudaGetLastError();
}

}  // extern "C"
