// Build of the uniform-grid occluder (see occluder.cuh): count -> scan -> fill, two C-ABI phases around one host
// read of the entry total.  Replaces optix_build_bvh (reference render/optixutils/c_src/torch_bindings.cpp:37-116).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"
#include "occluder.cuh"

using namespace gsb;

namespace {
constexpr int kThreads = 256;
inline int nblk(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

__global__ void k_params(const float* __restrict__ lo, const float* __restrict__ hi, int R, Occluder* occ,
                         const int32_t* cell_start, const unsigned long long* brick_occ) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-6f));
  float cell = ext * 1.0001f / (float)R;
  Occluder o;
  o.cell_start = cell_start;
  o.cell_tri_data = nullptr;
  o.brick_occ = brick_occ;
  o.cell_slabs = nullptr;
  o.nbx = o.nby = (R + 3) / 4;
  // cubic grid centred on the bounding box
  o.ox = 0.5f * (lo[0] + hi[0]) - 0.5f * cell * R;
  o.oy = 0.5f * (lo[1] + hi[1]) - 0.5f * cell * R;
  o.oz = 0.5f * (lo[2] + hi[2]) - 0.5f * cell * R;
  o.cell = cell;
  o.inv_cell = 1.f / cell;
  o.nx = o.ny = o.nz = R;
  *occ = o;
}

struct CellRange { int x0, x1, y0, y1, z0, z1; };

__device__ __forceinline__ CellRange tri_cells(const Occluder& o, float3 a, float3 b, float3 c) {
  const float pad = 1e-4f * o.cell;
  CellRange r;
  r.x0 = min(max((int)floorf((fminf(a.x, fminf(b.x, c.x)) - pad - o.ox) * o.inv_cell), 0), o.nx - 1);
  r.x1 = min(max((int)floorf((fmaxf(a.x, fmaxf(b.x, c.x)) + pad - o.ox) * o.inv_cell), 0), o.nx - 1);
  r.y0 = min(max((int)floorf((fminf(a.y, fminf(b.y, c.y)) - pad - o.oy) * o.inv_cell), 0), o.ny - 1);
  r.y1 = min(max((int)floorf((fmaxf(a.y, fmaxf(b.y, c.y)) + pad - o.oy) * o.inv_cell), 0), o.ny - 1);
  r.z0 = min(max((int)floorf((fminf(a.z, fminf(b.z, c.z)) - pad - o.oz) * o.inv_cell), 0), o.nz - 1);
  r.z1 = min(max((int)floorf((fmaxf(a.z, fmaxf(b.z, c.z)) + pad - o.oz) * o.inv_cell), 0), o.nz - 1);
  return r;
}

__device__ __forceinline__ float3 ldv(const float* v, int i) {
  return make_float3(__ldg(v + (size_t)i * 3), __ldg(v + (size_t)i * 3 + 1), __ldg(v + (size_t)i * 3 + 2));
}

// Separating-axis test triangle vs axis-aligned box (Akenine-Moeller): 3 box normals, the triangle normal, 9 edge cross
// products.  Vertices are given relative to the box centre; h = half extent (slightly padded by the caller).
__device__ __forceinline__ bool axis_separates(float ax, float ay, float az, float3 a, float3 b, float3 c, float h) {
  const float p0 = ax * a.x + ay * a.y + az * a.z, p1 = ax * b.x + ay * b.y + az * b.z, p2 = ax * c.x + ay * c.y + az * c.z;
  const float r = h * (fabsf(ax) + fabsf(ay) + fabsf(az));
  return fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r;
}
__device__ __forceinline__ bool tri_overlaps_box(float3 a, float3 b, float3 c, float h) {
  const float3 e0 = make_float3(b.x - a.x, b.y - a.y, b.z - a.z), e1 = make_float3(c.x - b.x, c.y - b.y, c.z - b.z),
               e2 = make_float3(a.x - c.x, a.y - c.y, a.z - c.z);
  // box normals
  if (axis_separates(1.f, 0.f, 0.f, a, b, c, h) || axis_separates(0.f, 1.f, 0.f, a, b, c, h) || axis_separates(0.f, 0.f, 1.f, a, b, c, h)) return false;
  // triangle normal
  if (axis_separates(e0.y * e1.z - e0.z * e1.y, e0.z * e1.x - e0.x * e1.z, e0.x * e1.y - e0.y * e1.x, a, b, c, h)) return false;
  // unit axes x edges
  const float3 e[3] = {e0, e1, e2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (axis_separates(0.f, -e[k].z, e[k].y, a, b, c, h)) return false;      // x cross e
    if (axis_separates(e[k].z, 0.f, -e[k].x, a, b, c, h)) return false;      // y cross e
    if (axis_separates(-e[k].y, e[k].x, 0.f, a, b, c, h)) return false;      // z cross e
  }
  return true;
}

template <bool FILL>
__global__ void __launch_bounds__(kThreads) k_bin(const float* __restrict__ verts, const int32_t* __restrict__ tris, int64_t F,
                                                  const Occluder* __restrict__ occ, int32_t* __restrict__ counts_or_cursor,
                                                  float4* __restrict__ cell_tri_data, uint32_t* __restrict__ cell_slabs) {
  int64_t f = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (f >= F) return;
  const Occluder o = *occ;
  const float3 a = ldv(verts, __ldg(tris + f * 3)), b = ldv(verts, __ldg(tris + f * 3 + 1)), c = ldv(verts, __ldg(tris + f * 3 + 2));
  // degenerate (zero-area) triangles can never be hit: skip them
  const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
  const float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
  if (nx == 0.f && ny == 0.f && nz == 0.f) return;
  const CellRange r = tri_cells(o, a, b, c);
  const float pad = 1e-3f * o.cell;
  for (int z = r.z0; z <= r.z1; ++z)
    for (int y = r.y0; y <= r.y1; ++y)
      for (int x = r.x0; x <= r.x1; ++x) {
        // exact overlap (not just the AABB): fewer (cell, triangle) entries => fewer wasted intersection tests per ray
        const float ccx = o.ox + (x + 0.5f) * o.cell, ccy = o.oy + (y + 0.5f) * o.cell, ccz = o.oz + (z + 0.5f) * o.cell;
        if (!tri_overlaps_box(make_float3(a.x - ccx, a.y - ccy, a.z - ccz), make_float3(b.x - ccx, b.y - ccy, b.z - ccz),
                              make_float3(c.x - ccx, c.y - ccy, c.z - ccz), 0.5f * o.cell * 1.001f))
          continue;
        const int cidx = (z * o.ny + y) * o.nx + x;
        if (FILL) {
          // slabs (eighths of the cell per axis) covered by the triangle's AABB clipped to this cell
          const float lx = o.ox + x * o.cell, ly = o.oy + y * o.cell, lz = o.oz + z * o.cell, s8 = 8.f * o.inv_cell;
          const int ax0 = min(max((int)floorf((fminf(a.x, fminf(b.x, c.x)) - pad - lx) * s8), 0), 7);
          const int ax1 = min(max((int)floorf((fmaxf(a.x, fmaxf(b.x, c.x)) + pad - lx) * s8), 0), 7);
          const int ay0 = min(max((int)floorf((fminf(a.y, fminf(b.y, c.y)) - pad - ly) * s8), 0), 7);
          const int ay1 = min(max((int)floorf((fmaxf(a.y, fmaxf(b.y, c.y)) + pad - ly) * s8), 0), 7);
          const int az0 = min(max((int)floorf((fminf(a.z, fminf(b.z, c.z)) - pad - lz) * s8), 0), 7);
          const int az1 = min(max((int)floorf((fmaxf(a.z, fmaxf(b.z, c.z)) + pad - lz) * s8), 0), 7);
          const unsigned m = (((2u << ax1) - (1u << ax0))) | (((2u << ay1) - (1u << ay0)) << 8) | (((2u << az1) - (1u << az0)) << 16);
          atomicOr(cell_slabs + cidx, m);
          const size_t e = 3 * (size_t)(__ldg(o.cell_start + cidx) + atomicAdd(counts_or_cursor + cidx, 1));
          cell_tri_data[e] = make_float4(a.x, a.y, a.z, ux);
          cell_tri_data[e + 1] = make_float4(uy, uz, vx, vy);
          cell_tri_data[e + 2] = make_float4(vz, 0.f, 0.f, 0.f);
        } else {
          atomicAdd(counts_or_cursor + cidx, 1);
        }
      }
}

// one thread per 4x4x4 brick: occupancy bits from the raw per-cell counts (before the scan turns them into offsets)
__global__ void __launch_bounds__(kThreads) k_brick_bits(const int32_t* __restrict__ counts, int R, int nb,
                                                         unsigned long long* __restrict__ bits) {
  const int b = blockIdx.x * kThreads + threadIdx.x;
  if (b >= nb * nb * nb) return;
  const int bx = b % nb, by = (b / nb) % nb, bz = b / (nb * nb);
  unsigned long long m = 0ull;
  for (int z = 0; z < 4; ++z)
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x) {
        const int cx = 4 * bx + x, cy = 4 * by + y, cz = 4 * bz + z;
        if (cx < R && cy < R && cz < R && counts[((size_t)cz * R + cy) * R + cx] > 0) m |= 1ull << ((z << 4) | (y << 2) | x);
      }
  bits[b] = m;
}

// ---- multi-block exclusive scan (in place) ----------------------------------------------------------------
constexpr int kScanTile = 2048;   // 256 threads x 8
__global__ void __launch_bounds__(kThreads) k_scan_tiles(int32_t* __restrict__ data, int64_t n, int32_t* __restrict__ tile_sums) {
  __shared__ int s_warp[kThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * 8;
  int v[8], sum = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] = (base + k < n) ? data[base + k] : 0; sum += v[k]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_warp[w];
  int run = woff + incl - sum;
#pragma unroll
  for (int k = 0; k < 8; ++k) { if (base + k < n) data[base + k] = run; run += v[k]; }
  if (threadIdx.x == kThreads - 1) tile_sums[blockIdx.x] = woff + incl;
}
__global__ void __launch_bounds__(1024) k_scan_sums(int32_t* __restrict__ sums, int n, int32_t* __restrict__ total) {
  __shared__ int warp_sums[32];
  __shared__ int chunk_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int carry = 0;
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    int x = i < n ? sums[i] : 0, incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
      warp_sums[lane] = wi - w;
      if (lane == 31) chunk_total = wi;
    }
    __syncthreads();
    if (i < n) sums[i] = carry + warp_sums[warp] + incl - x;
    carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(kThreads) k_scan_add(int32_t* __restrict__ data, int64_t n, const int32_t* __restrict__ tile_off,
                                                       const int32_t* __restrict__ total) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * 8;
  const int off = tile_off[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (base + k < n) data[base + k] += off;
  if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = *total;    // closing entry cell_start[ncells]
}

__global__ void k_set_entries(Occluder* occ, const float4* cell_tri_data, const uint32_t* cell_slabs) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { occ->cell_tri_data = cell_tri_data; occ->cell_slabs = cell_slabs; }
}


// ---- shadow-ray tracing: persistent warps over a compact ray list -------------------------------------------------
// list[2j] = (origin, ray id), list[2j+1] = (direction, -); vis[ray id] is pre-set to 1 and cleared on a hit.
// Lanes whose ray ended pick up new rays as soon as fewer than kRefill lanes of the warp are busy (Aila & Laine's
// persistent traversal with dynamic fetch): inline tracing inside the per-pixel sample loop kept only 2.4 of 32 lanes busy
// (ncu, profiles/r1c) because every sample waited for the slowest ray of the warp.  Build-time knobs (swept on the GPU with
// profiles/sweep_trace.sh, results in profiles/r1h_trace_kernel_ncu.md): refill threshold, look-ahead steps and triangle
// records per iteration, lanes that must wait before a cell entry is executed, CTA shape.
#ifndef GSB_TRACE_REFILL
#define GSB_TRACE_REFILL 24
#endif
#ifndef GSB_TRACE_BLOCKS
#define GSB_TRACE_BLOCKS 4
#endif
#ifndef GSB_TRACE_STEPS
#define GSB_TRACE_STEPS 4
#endif
constexpr int kRefill = GSB_TRACE_REFILL;
#ifndef GSB_TRACE_ENTER_VOTE
#define GSB_TRACE_ENTER_VOTE 4
#endif
constexpr int kEnterVote = GSB_TRACE_ENTER_VOTE;   // lanes that must wait for a cell entry before the warp executes it
#ifndef GSB_TRACE_BATCH
#define GSB_TRACE_BATCH 3
#endif
constexpr int kBatch = GSB_TRACE_BATCH;     // triangle records tested per iteration
constexpr int kSteps = GSB_TRACE_STEPS;      // cells the look-ahead DDA may advance per iteration
__device__ unsigned long long g_rays_traced = 0ull;     // running total, read by gsb_trace_ray_count (profiling aid)
#ifdef GSB_TRACE_STATS
__device__ unsigned long long g_trace_stats[4] = {0ull, 0ull, 0ull, 0ull};   // triangle tests, cell steps, occupied cells, hits
#define GSB_STAT(i) atomicAdd(&g_trace_stats[i], 1ull)
#else
#define GSB_STAT(i)
#endif
#ifndef GSB_TRACE_MIN_BLOCKS
#define GSB_TRACE_MIN_BLOCKS 4
#endif

// occupancy word of the 4x4x4 brick around cell (cx, cy, cz) and the cell's bit index inside it
__device__ __forceinline__ unsigned long long brick_word(const Occluder& g, int cx, int cy, int cz, int& bit) {
  bit = ((cz & 3) << 4) | ((cy & 3) << 2) | (cx & 3);
  return __ldg(g.brick_occ + ((cz >> 2) * g.nby + (cy >> 2)) * g.nbx + (cx >> 2));
}

// Sub-cell box test on entering an occupied cell: the triangles of a cell are often much smaller than the cell (on the
// random-SDF soup ~10 records per occupied cell), so a ray that misses the box spanned by the cell's slab bits skips them.
__device__ __forceinline__ bool ray_touches_slab_box(const Occluder& g, unsigned m, int cx, int cy, int cz, float ox, float oy,
                                                     float oz, float idx, float idy, float idz) {
  const float e = 0.125f * g.cell, pad = 2e-3f * g.cell;
  const float lx = g.ox + cx * g.cell - ox, ly = g.oy + cy * g.cell - oy, lz = g.oz + cz * g.cell - oz;
  const unsigned mx = m & 255u, my = (m >> 8) & 255u, mz = (m >> 16) & 255u;
  float a = (lx + (__ffs(mx) - 1) * e - pad) * idx, b = (lx + (32 - __clz(mx)) * e + pad) * idx;
  float tn = fminf(a, b), tf = fmaxf(a, b);
  a = (ly + (__ffs(my) - 1) * e - pad) * idy; b = (ly + (32 - __clz(my)) * e + pad) * idy;
  tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));
  a = (lz + (__ffs(mz) - 1) * e - pad) * idz; b = (lz + (32 - __clz(mz)) * e + pad) * idz;
  tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));
  return tn <= tf && tf >= 0.f;
}

#ifndef GSB_TRACE_THREADS
#define GSB_TRACE_THREADS 256
#endif
__global__ void __launch_bounds__(GSB_TRACE_THREADS, GSB_TRACE_MIN_BLOCKS) k_trace_list(const Occluder* __restrict__ occ_p, const float4* __restrict__ list,
                                                         const int32_t* __restrict__ count_p, int cap, int32_t* __restrict__ cursor,
                                                         uint8_t* __restrict__ vis) {
  const Occluder g = *occ_p;
  const int n = min(*count_p, cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_rays_traced, (unsigned long long)n);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const float gx1 = g.ox + g.nx * g.cell, gy1 = g.oy + g.ny * g.cell, gz1 = g.oz + g.nz * g.cell;
  const float big = 3.0e38f;
  // Software-pipelined traversal.  While a lane tests the triangles of its current cell (A, one per iteration) its DDA
  // already looks ahead for the next occupied cell on the occupancy bits (B, one step per iteration), so both code blocks
  // run with most lanes active; the triangle range and slab mask of the cell found are requested at discovery and consumed
  // at entry.  The rare entry itself (C: sub-box test, ~6 per ray against ~45 tests and ~50 steps) is executed only when
  // enough lanes wait for it: run divergently it kept 1.0 lane busy and took 32 % of all issue slots (ncu, profiles/r1g).
  bool have = false, exhausted = false, found = false, exited = false;
  int rid = 0;
  float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, tmx = 0, tmy = 0, tmz = 0, tdx = 0, tdy = 0, tdz = 0;
  int cx = 0, cy = 0, cz = 0, k0 = 0, k1 = 0, bit = 0, r0 = 0, r1 = 0;
  unsigned slab = 0u;
  unsigned long long bw = 0ull;
  for (;;) {
    const unsigned act = __ballot_sync(full, have);
    const int nact = __popc(act);
    if (!exhausted && nact < kRefill) {                             // warp-uniform refill
      const int nidle = 32 - nact;
      int base = 0;
      if (lane == 0) base = atomicAdd(cursor, nidle);
      base = __shfl_sync(full, base, 0);
      if (base + nidle >= n) exhausted = true;
      if (!have) {
        const int j = base + __popc(~act & ((1u << lane) - 1u));
        if (j < n) {
          const float4 a = __ldg(list + 2 * (size_t)j), b = __ldg(list + 2 * (size_t)j + 1);
          ox = a.x; oy = a.y; oz = a.z; rid = __float_as_int(a.w);
          dx = b.x; dy = b.y; dz = b.z;
          const float idx = 1.f / dx, idy = 1.f / dy, idz = 1.f / dz;
          float t0 = 0.f, t1 = big;
          const bool inside = ox >= g.ox && ox <= gx1 && oy >= g.oy && oy <= gy1 && oz >= g.oz && oz <= gz1;
          if (!inside) {
            float u = (g.ox - ox) * idx, v = (gx1 - ox) * idx;
            if (dx == 0.f) { if (ox < g.ox || ox > gx1) t1 = -1.f; } else { t0 = fmaxf(t0, fminf(u, v)); t1 = fminf(t1, fmaxf(u, v)); }
            u = (g.oy - oy) * idy; v = (gy1 - oy) * idy;
            if (dy == 0.f) { if (oy < g.oy || oy > gy1) t1 = -1.f; } else { t0 = fmaxf(t0, fminf(u, v)); t1 = fminf(t1, fmaxf(u, v)); }
            u = (g.oz - oz) * idz; v = (gz1 - oz) * idz;
            if (dz == 0.f) { if (oz < g.oz || oz > gz1) t1 = -1.f; } else { t0 = fmaxf(t0, fminf(u, v)); t1 = fminf(t1, fmaxf(u, v)); }
          }
          if (t0 <= t1) {
            const float ex = ox + dx * t0, ey = oy + dy * t0, ez = oz + dz * t0;
            cx = min(max((int)floorf((ex - g.ox) * g.inv_cell), 0), g.nx - 1);
            cy = min(max((int)floorf((ey - g.oy) * g.inv_cell), 0), g.ny - 1);
            cz = min(max((int)floorf((ez - g.oz) * g.inv_cell), 0), g.nz - 1);
            tmx = dx != 0.f ? (g.ox + (cx + (dx > 0.f ? 1 : 0)) * g.cell - ox) * idx : big;
            tmy = dy != 0.f ? (g.oy + (cy + (dy > 0.f ? 1 : 0)) * g.cell - oy) * idy : big;
            tmz = dz != 0.f ? (g.oz + (cz + (dz > 0.f ? 1 : 0)) * g.cell - oz) * idz : big;
            tdx = dx != 0.f ? g.cell * fabsf(idx) : big;
            tdy = dy != 0.f ? g.cell * fabsf(idy) : big;
            tdz = dz != 0.f ? g.cell * fabsf(idz) : big;
            k0 = k1 = 0;
            exited = false;
            bw = brick_word(g, cx, cy, cz, bit);
            found = (bw >> bit) & 1ull;
            if (found) {
              const int c = (cz * g.ny + cy) * g.nx + cx;
              slab = __ldg(g.cell_slabs + c); r0 = __ldg(g.cell_start + c); r1 = __ldg(g.cell_start + c + 1);
            }
            have = true;
          }
        }
      }
    }
    if (exhausted && __ballot_sync(full, have) == 0u) break;
    // ---- A: kBatch triangles of the current cell (branch-free tests) ----
    // (measured and dropped, profiles/r1h: prefetching the next record into registers (78 regs, 3 CTAs/SM: 97 ms vs 77),
    //  prefetch.global.L1 of the next record (88 ms), issuing the record loads before B (no change))
    if (have && k0 < k1) {
      const float4* td = g.cell_tri_data + (size_t)k0 * 3;
      // kBatch records per iteration: their loads are in flight together, so one exposed HBM/L2 round trip serves kBatch
      // tests (1 -> 2 records: 76 -> 63 ms on the N=103 probe); records past the end of the cell repeat the last one
      float4 ra[kBatch], rb[kBatch];
      float rc[kBatch];
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const float4* t = td + 3 * min(q, k1 - k0 - 1);
        ra[q] = __ldg(t); rb[q] = __ldg(t + 1); rc[q] = __ldg(reinterpret_cast<const float*>(t + 2));
      }
      bool hit = false;
#pragma unroll
      for (int q = 0; q < kBatch; ++q) hit |= ray_hits_triangle_bf(ra[q], rb[q], rc[q], ox, oy, oz, dx, dy, dz);
      k0 += kBatch;
      GSB_STAT(0);
      if (hit) {
        vis[rid] = 0;
        have = false;
        GSB_STAT(3);
      }
    }
    // ---- B: look ahead for the next occupied cell ----
    if (have && !found && !exited) {
#pragma unroll 1
      for (int s = 0; s < kSteps; ++s) {
        // one DDA step; the occupancy word stays in registers while the ray is inside the brick and the bit index moves with
        // the step, so an empty cell costs ~20 instructions and no memory access
        // (written with selects: as three if-branches the compiler emitted real branches and the lanes of a warp split
        //  three ways on every step)
        const bool ax = tmx <= tmy && tmx <= tmz, ay = !ax && tmy <= tmz;
        const int sg = (ax ? dx : (ay ? dy : dz)) > 0.f ? 1 : -1;
        cx += ax ? sg : 0;
        cy += ay ? sg : 0;
        cz += (ax || ay) ? 0 : sg;
        tmx += ax ? tdx : 0.f;
        tmy += ay ? tdy : 0.f;
        tmz += (ax || ay) ? 0.f : tdz;
        bit += sg << (ax ? 0 : (ay ? 2 : 4));
        const bool crossed = ((ax ? cx : (ay ? cy : cz)) & 3) == (sg > 0 ? 0 : 3);   // left the current 4x4x4 brick?
        GSB_STAT(1);
        if (crossed) {
          // the grid is left through a brick face: cells past nx inside the last brick exist as empty cells
          if ((unsigned)(cx >> 2) >= (unsigned)g.nbx || (unsigned)(cy >> 2) >= (unsigned)g.nby || (unsigned)(cz >> 2) >= (unsigned)g.nbx) {
            exited = true;
            break;
          }
          bw = brick_word(g, cx, cy, cz, bit);
        }
        if ((bw >> bit) & 1ull) {
          const int c = (cz * g.ny + cy) * g.nx + cx;
          slab = __ldg(g.cell_slabs + c); r0 = __ldg(g.cell_start + c); r1 = __ldg(g.cell_start + c + 1);
          found = true;
          break;
        }
      }
    }
    // ---- C: current cell finished -> leave the grid, or enter the cell found by the look-ahead (voted) ----
    const bool drained = have && k0 >= k1;
    if (drained && exited) have = false;                            // no hit anywhere: stays visible
    const bool wants_entry = drained && found;
    const int n_entry = __popc(__ballot_sync(full, wants_entry));
    const int n_busy = __popc(__ballot_sync(full, have && !wants_entry));
    if (n_entry > 0 && (n_entry >= kEnterVote || n_entry >= n_busy)) {
      if (wants_entry) {
        // 1/d from the DDA increments (td = cell / |d|); +-inf for axis-parallel rays, handled by fmin/fmax
        const bool touch = ray_touches_slab_box(g, slab, cx, cy, cz, ox, oy, oz, copysignf(tdx * g.inv_cell, dx),
                                                copysignf(tdy * g.inv_cell, dy), copysignf(tdz * g.inv_cell, dz));
        GSB_STAT(2);
        k0 = touch ? r0 : 0;
        k1 = touch ? r1 : 0;
        found = false;


      }
    }
  }
}

}  // namespace

extern "C" {

size_t gsb_occluder_struct_bytes(void) { return sizeof(Occluder); }

int64_t gsb_occluder_brick_words(int grid_res) { const int64_t nb = (grid_res + 3) / 4; return nb * nb * nb; }

int64_t gsb_occluder_scan_ws_ints(int64_t n_cells) { return (n_cells + kScanTile - 1) / kScanTile + 1; }

int gsb_occluder_build_count(const float* verts, const int32_t* tris, int64_t n_faces, const float* bounds_lo,
                             const float* bounds_hi, int grid_res, void* occluder, int32_t* cell_start, int32_t* scan_ws,
                             uint64_t* brick_bits, int32_t* total, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (grid_res < 1 || grid_res > 1024) return (int)cudaErrorInvalidValue;
  const int64_t n_cells = (int64_t)grid_res * grid_res * grid_res;
  cudaError_t e = cudaMemsetAsync(cell_start, 0, sizeof(int32_t) * (size_t)(n_cells + 1), stream);
  if (e != cudaSuccess) return (int)e;
  k_params<<<1, 32, 0, stream>>>(bounds_lo, bounds_hi, grid_res, (Occluder*)occluder, cell_start,
                                 (const unsigned long long*)brick_bits);
  if (n_faces > 0)
    k_bin<false><<<nblk(n_faces), kThreads, 0, stream>>>(verts, tris, n_faces, (const Occluder*)occluder, cell_start, nullptr, nullptr);
  const int nb = (grid_res + 3) / 4;
  k_brick_bits<<<nblk((int64_t)nb * nb * nb), kThreads, 0, stream>>>(cell_start, grid_res, nb, (unsigned long long*)brick_bits);
  const int n_tiles = (int)((n_cells + kScanTile - 1) / kScanTile);
  k_scan_tiles<<<n_tiles, kThreads, 0, stream>>>(cell_start, n_cells, scan_ws);
  k_scan_sums<<<1, 1024, 0, stream>>>(scan_ws, n_tiles, total);
  k_scan_add<<<n_tiles, kThreads, 0, stream>>>(cell_start, n_cells, scan_ws, total);
  return (int)cudaGetLastError();
}

int gsb_occluder_build_fill(const float* verts, const int32_t* tris, int64_t n_faces, int grid_res, void* occluder,
                            int32_t* cursor, uint32_t* cell_slabs, float* cell_tri_data, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t n_cells = (int64_t)grid_res * grid_res * grid_res;
  cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)n_cells, stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(cell_slabs, 0, sizeof(uint32_t) * (size_t)n_cells, stream);
  if (e != cudaSuccess) return (int)e;
  k_set_entries<<<1, 32, 0, stream>>>((Occluder*)occluder, (const float4*)cell_tri_data, cell_slabs);
  if (n_faces > 0)
    k_bin<true><<<nblk(n_faces), kThreads, 0, stream>>>(verts, tris, n_faces, (const Occluder*)occluder, cursor,
                                                        (float4*)cell_tri_data, cell_slabs);
  return (int)cudaGetLastError();
}

int gsb_trace_shadow_rays(const void* occluder, const void* ray_list, const int32_t* ray_count, int64_t ray_cap,
                          int32_t* fetch_counter, uint8_t* vis, void* stream_) {
  // persistent grid: 4 CTAs of 256 threads per SM (61 registers/thread)
  k_trace_list<<<148 * (GSB_TRACE_BLOCKS > GSB_TRACE_MIN_BLOCKS ? GSB_TRACE_BLOCKS : GSB_TRACE_MIN_BLOCKS), GSB_TRACE_THREADS, 0, (cudaStream_t)stream_>>>((const Occluder*)occluder, (const float4*)ray_list, ray_count,
                                                               (int)(ray_cap < 0x7fffffff ? ray_cap : 0x7fffffff),
                                                               fetch_counter, vis);
  return (int)cudaGetLastError();
}

/* Traversal counters {triangle tests, cell steps, occupied cells entered, hits}; all zero unless the library was built with
 * -DGSB_TRACE_STATS (profiling builds only: the counters are global atomics). */
void gsb_trace_stats(uint64_t* out4, int reset) {
#ifdef GSB_TRACE_STATS
  unsigned long long v[4];
  cudaMemcpyFromSymbol(v, g_trace_stats, sizeof(v));
  for (int i = 0; i < 4; ++i) out4[i] = v[i];
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(g_trace_stats, z, sizeof(z)); }
#else
  (void)reset;
  for (int i = 0; i < 4; ++i) out4[i] = 0;
#endif
}

/* Rays handed to the trace kernel since the last reset (profiling aid; synchronises the device). */
uint64_t gsb_trace_ray_count(int reset) {
  unsigned long long v = 0ull;
  cudaMemcpyFromSymbol(&v, g_rays_traced, sizeof(v));
  if (reset) {
    unsigned long long z = 0ull;
    cudaMemcpyToSymbol(g_rays_traced, &z, sizeof(z));
  }
  return (uint64_t)v;
}

}  // extern "C"
