// Occluder for shadow rays: build (count -> scan -> fill, two C-ABI phases around one host read of the entry total) and
// the persistent any-hit trace kernel over a compact ray list.  Replaces optix_build_bvh + optixTrace of the reference
// (render/optixutils/c_src/torch_bindings.cpp:37-116, envsampling/kernel.cu:101-118).  Structure and traversal: trace_core.cuh.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"
#include "trace_core.cuh"

using namespace gsb;

namespace {
constexpr int kThreads = 256;
inline int nblk(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

__global__ void k_params(const float* __restrict__ lo, const float* __restrict__ hi, int R, OccGrid* occ,
                         const unsigned long long* brick_occ) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-6f));
  float cell = ext * 1.0001f / (float)R;
  OccGrid o;
  o.brick_occ = brick_occ;
  o.cell_rec = nullptr;
  o.tri_rec = nullptr;
  // cubic grid centred on the bounding box
  o.ox = 0.5f * (lo[0] + hi[0]) - 0.5f * cell * R;
  o.oy = 0.5f * (lo[1] + hi[1]) - 0.5f * cell * R;
  o.oz = 0.5f * (lo[2] + hi[2]) - 0.5f * cell * R;
  o.cell = cell;
  o.inv_cell = 1.f / cell;
  o.n = R;
  o.nb = (R + 3) / 4;
  *occ = o;
}

struct CellRange { int x0, x1, y0, y1, z0, z1; };

__device__ __forceinline__ CellRange tri_cells(const OccGrid& o, float3 a, float3 b, float3 c) {
  const float pad = 1e-4f * o.cell;
  CellRange r;
  r.x0 = min(max((int)floorf((fminf(a.x, fminf(b.x, c.x)) - pad - o.ox) * o.inv_cell), 0), o.n - 1);
  r.x1 = min(max((int)floorf((fmaxf(a.x, fmaxf(b.x, c.x)) + pad - o.ox) * o.inv_cell), 0), o.n - 1);
  r.y0 = min(max((int)floorf((fminf(a.y, fminf(b.y, c.y)) - pad - o.oy) * o.inv_cell), 0), o.n - 1);
  r.y1 = min(max((int)floorf((fmaxf(a.y, fmaxf(b.y, c.y)) + pad - o.oy) * o.inv_cell), 0), o.n - 1);
  r.z0 = min(max((int)floorf((fminf(a.z, fminf(b.z, c.z)) - pad - o.oz) * o.inv_cell), 0), o.n - 1);
  r.z1 = min(max((int)floorf((fmaxf(a.z, fmaxf(b.z, c.z)) + pad - o.oz) * o.inv_cell), 0), o.n - 1);
  return r;
}

__device__ __forceinline__ float3 ldv(const float* v, int i) {
  return make_float3(__ldg(v + (size_t)i * 3), __ldg(v + (size_t)i * 3 + 1), __ldg(v + (size_t)i * 3 + 2));
}

// One thread per triangle: every cell whose (slightly grown) box the triangle really overlaps (exact SAT, not just the
// AABB: fewer (cell, triangle) entries => fewer wasted intersection tests per ray) gets an entry.  FILL writes the record and
// leaves (cell id, triangle id) in its two spare words for k_subvoxels.
template <bool FILL>
__global__ void __launch_bounds__(kThreads) k_bin(const float* __restrict__ verts, const int32_t* __restrict__ tris, int64_t F,
                                                  const OccGrid* __restrict__ occ, int32_t* __restrict__ counts_or_cursor,
                                                  float4* __restrict__ tri_rec, uint4* __restrict__ cell_rec) {
  int64_t f = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (f >= F) return;
  const OccGrid o = *occ;
  const float3 a = ldv(verts, __ldg(tris + f * 3)), b = ldv(verts, __ldg(tris + f * 3 + 1)), c = ldv(verts, __ldg(tris + f * 3 + 2));
  // degenerate (zero-area) triangles can never be hit: skip them
  const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
  const float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
  if (nx == 0.f && ny == 0.f && nz == 0.f) return;
  const CellRange r = tri_cells(o, a, b, c);
  for (int z = r.z0; z <= r.z1; ++z)
    for (int y = r.y0; y <= r.y1; ++y)
      for (int x = r.x0; x <= r.x1; ++x) {
        const float ccx = o.ox + (x + 0.5f) * o.cell, ccy = o.oy + (y + 0.5f) * o.cell, ccz = o.oz + (z + 0.5f) * o.cell;
        if (!tri_overlaps_box(make_float3(a.x - ccx, a.y - ccy, a.z - ccz), make_float3(b.x - ccx, b.y - ccy, b.z - ccz),
                              make_float3(c.x - ccx, c.y - ccy, c.z - ccz), 0.5f * o.cell * 1.001f))
          continue;
        const int64_t cid = cell_id(x, y, z, o.nb);
        if (FILL) {
          const size_t e = 3 * (size_t)(cell_rec[cid].x + (uint32_t)atomicAdd(counts_or_cursor + cid, 1));
          tri_rec[e] = make_float4(a.x, a.y, a.z, ux);
          tri_rec[e + 1] = make_float4(uy, uz, vx, vy);
          tri_rec[e + 2] = make_float4(vz, __int_as_float((int)cid), __int_as_float((int)f), 0.f);
        } else {
          atomicAdd(counts_or_cursor + cid, 1);
        }
      }
}

// Sub-voxel bits: four threads per (cell, triangle) entry, one z-slice of the cell's 4x4x4 sub-voxels each (up to 16 exact box
// tests), OR-ed into the cell record.  (Inside k_bin<FILL> -- one thread walking all the cells of its triangle and all their
// sub-voxels -- this was 10.1 of the 10.7 ms of the whole build, replicated on every rank.)
__global__ void __launch_bounds__(kThreads) k_subvoxels(const float* __restrict__ verts, const int32_t* __restrict__ tris,
                                                        const OccGrid* __restrict__ occ, const int32_t* __restrict__ n_entries,
                                                        const float4* __restrict__ tri_rec, uint4* __restrict__ cell_rec) {
  const OccGrid o = *occ;
  const int64_t n = (int64_t)(*n_entries) * 4;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const int64_t e = i >> 2;
    const int sz = (int)(i & 3);
    const float4 tail = __ldg(tri_rec + 3 * e + 2);
    const int cid = __float_as_int(tail.y);
    const int64_t f = __float_as_int(tail.z);
    const float3 a = ldv(verts, __ldg(tris + f * 3)), b = ldv(verts, __ldg(tris + f * 3 + 1)), c = ldv(verts, __ldg(tris + f * 3 + 2));
    const int brick = cid >> 6, bx = brick % o.nb, by = (brick / o.nb) % o.nb, bz = brick / (o.nb * o.nb);
    const int x = 4 * bx + (cid & 3), y = 4 * by + ((cid >> 2) & 3), z = 4 * bz + ((cid >> 4) & 3);
    const uint32_t m = subvoxel_slice(a, b, c, o.ox + x * o.cell, o.oy + y * o.cell, o.oz + z * o.cell, o.cell, sz);
    if (m != 0u) atomicOr((sz < 2 ? &cell_rec[cid].z : &cell_rec[cid].w), m << (16 * (sz & 1)));
  }
}

// one thread per cell (brick-major order: 64 consecutive cells = one brick = two warps): occupancy bits by ballot, from the
// raw per-cell counts (before the scan turns them into offsets)
__global__ void __launch_bounds__(kThreads) k_brick_bits(const int32_t* __restrict__ counts, int64_t n_cells, uint32_t* __restrict__ bits32) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;      // n_cells is a multiple of 64: whole warps
  const unsigned m = __ballot_sync(0xffffffffu, i < n_cells && counts[i] > 0);
  if ((threadIdx.x & 31) == 0 && i < n_cells) bits32[i >> 5] = m;
}

// cell records {first entry, entries, empty sub-voxel mask} from the scanned counts
__global__ void __launch_bounds__(kThreads) k_cell_recs(const int32_t* __restrict__ start, int64_t n_cells, uint4* __restrict__ rec) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_cells) return;
  const int s0 = start[i], s1 = start[i + 1];
  rec[i] = make_uint4((uint32_t)s0, (uint32_t)(s1 - s0), 0u, 0u);
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


__global__ void k_set_tables(OccGrid* occ, const uint4* cell_rec, const float4* tri_rec) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { occ->cell_rec = cell_rec; occ->tri_rec = tri_rec; }
}

// ---- shadow-ray tracing: persistent warps over a compact ray list -------------------------------------------------
// list[2j] = (origin, ray id), list[2j+1] = (direction, -); vis[ray id] is pre-set to 1 and cleared on a hit.
// A ray is in one of three states:
//   SEARCH  kSteps cell steps on the brick bits (trace_core.cuh: no memory access inside a brick, one 8-byte load per brick
//           crossed)                                                               -> DESC at an occupied cell, or leaves the grid
//   DESC    fetch the 16-byte cell record and walk the cell's sub-voxel bits       -> TEST at an occupied sub-voxel, else SEARCH
//   TEST    kRounds x kBatch triangle records of the cell (the loads of a round are in flight together)
//                                                                                  -> hit: ray done; list end: SEARCH
// A block of code executed for one lane costs the warp as much as for 32, and the three kinds of work alternate per ray every few
// hundred instructions.  History of the schedule (profiles/r2_trace_sweeps.md): one ray per lane, blocks gated by votes:
// 16 / 8 / 11 of 32 lanes busy; two ray contexts per lane in shared memory, each block picking one of the lane's own contexts:
// 10 / 8 / 11 lanes at 510 warp instructions per ray.  Now the contexts of a warp form ONE POOL of 32 * kPoolK slots in shared
// memory ([field][slot]) that any lane may work on: every trip around the loop the warp counts the slots per state (one REDUX),
// picks the state that fills the most lanes, hands the i-th slot in that state to lane i (rank by ballot, through a 32-byte
// list) and runs only that block -- measured 27 / 24 / 25 lanes per block.  Free slots are refilled from the ray list with a
// warp-aggregated fetch (persistent threads).  The grid description travels as a kernel parameter (constant bank).
#ifndef GSB_TRACE_POOL
#define GSB_TRACE_POOL 2             // slots per lane.  3 slots x 6 CTAs: 53 ms against 39 (fewer warps, less L1)
#endif
#ifndef GSB_TRACE_POOL_THREADS
#define GSB_TRACE_POOL_THREADS 128
#endif
#ifndef GSB_TRACE_POOL_BLOCKS
#define GSB_TRACE_POOL_BLOCKS 8      // CTAs per SM: what the pool's shared memory allows; 7 leaves more L1 and is 5 % slower
#endif
#ifndef GSB_TRACE_POOL_REFILL
#define GSB_TRACE_POOL_REFILL 16     // free slots before the warp fetches rays
#endif
#ifndef GSB_TRACE_POOL_BIAS_T
#define GSB_TRACE_POOL_BIAS_T 0      // lanes of head start for TEST / DESC over SEARCH when the block is chosen
#endif
#ifndef GSB_TRACE_POOL_BIAS_D
#define GSB_TRACE_POOL_BIAS_D 0
#endif
#ifndef GSB_TRACE_POOL_TROUNDS
#define GSB_TRACE_POOL_TROUNDS 2     // batches of triangle records per TEST execution (lanes that finish early idle for the rest)
#endif
#ifndef GSB_TRACE_STEPS
#define GSB_TRACE_STEPS 5
#endif
#ifndef GSB_TRACE_BATCH
#define GSB_TRACE_BATCH 4
#endif
#ifndef GSB_TRACE_RAY_BLOCK
#define GSB_TRACE_RAY_BLOCK 256      // consecutive rays a warp takes from the list per counter update (32-512: same time; 2048: +4 %, 8192: +14 %)
#endif
constexpr int kRayBlock = GSB_TRACE_RAY_BLOCK;
#ifndef GSB_TRACE_FINE_CAP
#define GSB_TRACE_FINE_CAP 0         // sub-voxel steps per DESC execution (0: the walk always runs to its end)
#endif
constexpr int kFineCap = GSB_TRACE_FINE_CAP;
constexpr int kSteps = GSB_TRACE_STEPS;            // cell steps per SEARCH execution
constexpr int kBatch = GSB_TRACE_BATCH;            // triangle records per TEST round
__device__ unsigned long long g_rays_traced = 0ull;     // running total, read by gsb_trace_ray_count (profiling aid)
#ifdef GSB_TRACE_STATS
__device__ unsigned long long g_trace_stats[16] = {0ull};
#define GSB_STAT(i, n) atomicAdd(&g_trace_stats[i], (unsigned long long)(n))
#else
#define GSB_STAT(i, n)
#endif
enum { ST_SEARCH = 0, ST_DESC = 1, ST_TEST = 2, ST_IDLE = 3 };
enum { F_TMX, F_TMY, F_TMZ, F_TDX, F_TDY, F_TDZ, F_T0, F_POS, F_WLO, F_WHI, F_FLIP, F_BLIN,
       F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_RID, F_K0, F_K1, F_COUNT };
constexpr int kPoolK = GSB_TRACE_POOL;
constexpr int kPoolSlots = 32 * kPoolK;
constexpr int kPoolThreads = GSB_TRACE_POOL_THREADS;
constexpr int kPoolWarpWords = (F_COUNT * kPoolSlots + kPoolSlots / 4 + 8 + 31) / 32 * 32;   // contexts + state bytes + 32-byte list
constexpr size_t kPoolSmemBytes = (size_t)(kPoolThreads / 32) * kPoolWarpWords * sizeof(uint32_t);

__global__ void __launch_bounds__(GSB_TRACE_POOL_THREADS, GSB_TRACE_POOL_BLOCKS) k_trace_pool(const __grid_constant__ OccGrid g, const float4* __restrict__ list,
                                                        const int32_t* __restrict__ count_p, int cap, int32_t* __restrict__ cursor,
                                                        uint8_t* __restrict__ vis) {
  extern __shared__ uint32_t ctx_smem[];
  uint32_t* const cx = ctx_smem + (threadIdx.x >> 5) * kPoolWarpWords;
  float* const cf = reinterpret_cast<float*>(cx);
  uint8_t* const stt = reinterpret_cast<uint8_t*>(cx + F_COUNT * kPoolSlots);     // state of every slot
  uint8_t* const sel = stt + kPoolSlots;                                         // slot handed to lane i this trip
#define PU(f) cx[(f) * kPoolSlots + slot]
#define PF(f) cf[(f) * kPoolSlots + slot]
  const int n = min(*count_p, cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_rays_traced, (unsigned long long)n);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  bool exhausted = false;
  int wbase = 0, wend = 0;                 // this warp's block of the ray list
#pragma unroll
  for (int k = 0; k < kPoolK; ++k) stt[lane + 32 * k] = (uint8_t)ST_IDLE;
  for (;;) {
    __syncwarp();
    // ---- census: slots per state (ST_SEARCH / DESC / TEST / IDLE = byte 0 / 1 / 2 / 3 of one packed sum) ----
    uint32_t st[kPoolK], packed = 0u;
#pragma unroll
    for (int k = 0; k < kPoolK; ++k) {
      st[k] = stt[lane + 32 * k];
      packed += 1u << (8u * st[k]);
    }
    const uint32_t tot = __reduce_add_sync(full, packed);
    const int n_search = (int)(tot & 255u), n_desc = (int)((tot >> 8) & 255u), n_test = (int)((tot >> 16) & 255u), n_idle = (int)(tot >> 24);
    uint32_t pick;
    int n_pick;
    if (!exhausted && n_idle >= GSB_TRACE_POOL_REFILL) {
      pick = ST_IDLE; n_pick = n_idle;
    } else {
      if (n_search + n_desc + n_test == 0) break;                 // pool empty and no rays left (not exhausted => n_idle is the pool)
      const int e_s = min(n_search, 32), e_d = min(n_desc, 32) + (n_desc ? GSB_TRACE_POOL_BIAS_D : 0), e_t = min(n_test, 32) + (n_test ? GSB_TRACE_POOL_BIAS_T : 0);
      if (e_t >= e_d && e_t >= e_s) { pick = ST_TEST; n_pick = n_test; }
      else if (e_d >= e_s) { pick = ST_DESC; n_pick = n_desc; }
      else { pick = ST_SEARCH; n_pick = n_search; }
    }
    // ---- lane i takes the i-th slot that is in the picked state ----
    {
      int base = 0;
#pragma unroll
      for (int k = 0; k < kPoolK; ++k) {
        const unsigned m = __ballot_sync(full, st[k] == pick);
        const int r = base + __popc(m & lt);
        if (st[k] == pick && r < 32) sel[r] = (uint8_t)(lane + 32 * k);
        base += __popc(m);
      }
    }
    __syncwarp();
    const int n_act = min(n_pick, 32);
    const bool act = lane < n_act;
    const int slot = act ? (int)sel[lane] : 0;
#ifdef GSB_TRACE_STATS
    if (lane == 0) { GSB_STAT(8 + 2 * pick, 1); GSB_STAT(9 + 2 * pick, n_act); }
#endif
    if (pick == ST_IDLE) {
      // ---- refill: one new ray per free slot.  A warp takes the list in blocks of kRayBlock consecutive rays (one counter update
      // per block; consecutive rays come from neighbouring pixels.  Measured neutral for L1: profiles/r2_trace_sweeps.md U) ----
      if (wbase >= wend) {
        int b = 0;
        if (lane == 0) b = atomicAdd(cursor, kRayBlock);
        b = __shfl_sync(full, b, 0);
        wbase = min(b, n);
        wend = min(b + kRayBlock, n);
        if (wbase >= wend) exhausted = true;
      }
      const int n_take = min(n_act, wend - wbase);
      const int base = wbase;
      wbase += n_take;
      const int j = base + lane;
      if (lane < n_take) {
        const float4 a = __ldg(list + 2 * (size_t)j), b = __ldg(list + 2 * (size_t)j + 1);
        Trav s;
        if (trav_setup(s, g, a.x, a.y, a.z, b.x, b.y, b.z)) {
          PF(F_TMX) = s.tmx; PF(F_TMY) = s.tmy; PF(F_TMZ) = s.tmz;
          PF(F_TDX) = s.tdx; PF(F_TDY) = s.tdy; PF(F_TDZ) = s.tdz;
          PF(F_T0) = s.t0;
          PU(F_POS) = s.pos; PU(F_WLO) = s.wlo; PU(F_WHI) = s.whi; PU(F_FLIP) = s.flip; PU(F_BLIN) = (uint32_t)s.blin;
          PF(F_OX) = a.x; PF(F_OY) = a.y; PF(F_OZ) = a.z;
          PF(F_DX) = b.x; PF(F_DY) = b.y; PF(F_DZ) = b.z;
          PU(F_RID) = (uint32_t)__float_as_int(a.w);
          PU(F_K1) = 0u;
          stt[slot] = (uint8_t)(trav_bit(s) ? ST_DESC : ST_SEARCH);
        }
      }
    } else if (pick == ST_SEARCH) {
      // ---- SEARCH: kSteps cell steps ----
      if (act) {
        Trav s;
        s.tmx = PF(F_TMX); s.tmy = PF(F_TMY); s.tmz = PF(F_TMZ);
        s.tdx = PF(F_TDX); s.tdy = PF(F_TDY); s.tdz = PF(F_TDZ);
        s.t0 = 0.f;
        s.pos = PU(F_POS); s.wlo = PU(F_WLO); s.whi = PU(F_WHI); s.flip = PU(F_FLIP); s.blin = (int32_t)PU(F_BLIN);
        trav_strides(s, g);
        int r = TR_CONT;
#pragma unroll
        for (int i = 0; i < kSteps; ++i) {
          if (r == TR_CONT) {
            GSB_STAT(1, 1);
            r = trav_step(s, g);
          }
        }
        PF(F_TMX) = s.tmx; PF(F_TMY) = s.tmy; PF(F_TMZ) = s.tmz;
        PU(F_POS) = s.pos; PU(F_WLO) = s.wlo; PU(F_WHI) = s.whi; PU(F_BLIN) = (uint32_t)s.blin;
        if (r != TR_CONT) stt[slot] = (uint8_t)(r == TR_EXIT ? ST_IDLE : ST_DESC);       // left the grid: the ray stays visible
      }
    } else if (pick == ST_DESC) {
      // ---- DESC: enter an occupied cell, walk its sub-voxel bits ----
      if (act) {
        Trav s;
        s.tmx = PF(F_TMX); s.tmy = PF(F_TMY); s.tmz = PF(F_TMZ);
        s.tdx = PF(F_TDX); s.tdy = PF(F_TDY); s.tdz = PF(F_TDZ);
        s.t0 = PF(F_T0);
        s.pos = PU(F_POS); s.flip = PU(F_FLIP); s.blin = (int32_t)PU(F_BLIN);
        s.wlo = s.whi = 0u;
        s.sx = s.sy = s.sz = 0;
        // the sub-voxel walk is bounded per execution (the loop runs until the slowest lane is through): an unfinished walk keeps
        // its position in F_K0, marked by F_K1 = ~0, and the slot stays in DESC
        const uint32_t k1 = PU(F_K1);
        const bool resume = kFineCap > 0 && k1 == 0xffffffffu;
        const uint4 rec = GSB_LDG_REC(g.cell_rec + trav_cell(s));
        Fine f;
        if (resume) fine_resume(s, PU(F_K0), f);
        else fine_enter(s, g, PF(F_DX), PF(F_DY), PF(F_DZ), f);
        uint32_t fine_steps;
        const int r = fine_walk(rec.z, rec.w, s.flip, f, kFineCap > 0 ? kFineCap : (1 << 30), fine_steps);
        if (!resume) GSB_STAT(2, 1);
        GSB_STAT(4, fine_steps);
        if (r == FINE_MORE) {
          PU(F_K0) = f.b; PU(F_K1) = 0xffffffffu;
        } else {
          PU(F_K0) = rec.x; PU(F_K1) = rec.x + rec.y;          // (a cell list never ends at entry 2^32 - 1: F_K1 != ~0)
          stt[slot] = (uint8_t)(r == FINE_HIT ? ST_TEST : ST_SEARCH);
          if (r == FINE_HIT) GSB_STAT(5, 1);
        }
      }
    } else {
      // ---- TEST: up to kRounds x kBatch triangle records of the cell ----
      bool busy = act;
      uint32_t k0 = 0u, k1 = 0u;
      float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
      if (act) {
        k0 = PU(F_K0); k1 = PU(F_K1);
        ox = PF(F_OX); oy = PF(F_OY); oz = PF(F_OZ); dx = PF(F_DX); dy = PF(F_DY); dz = PF(F_DZ);
      }
#pragma unroll 1
      for (int round = 0; round < GSB_TRACE_POOL_TROUNDS; ++round) {
        if (busy) {
          const float4* td = g.tri_rec + (size_t)k0 * 3;
          float4 ra[kBatch], rb[kBatch];
          float rc[kBatch];
#pragma unroll
          for (int q = 0; q < kBatch; ++q) {                          // records past the end of the cell repeat the last one
            const float4* t = td + 3 * min((uint32_t)q, k1 - k0 - 1u);
            ra[q] = GSB_LDG_TRI(t); rb[q] = GSB_LDG_TRI(t + 1); rc[q] = GSB_LDG_TRI(reinterpret_cast<const float*>(t + 2));
          }
          bool hit = false;
#pragma unroll
          for (int q = 0; q < kBatch; ++q) hit |= ray_hits_triangle(ra[q], rb[q], rc[q], ox, oy, oz, dx, dy, dz);
          GSB_STAT(0, min((uint32_t)kBatch, k1 - k0));
#ifdef GSB_TRACE_STATS
          {
            // analysis only (profiling builds): how many of these tests a PER-TRIANGLE sub-voxel mask would have kept -- counter 6:
            // the triangle's sub-voxels meet the sub-voxels the ray crosses in this cell; counter 7: they contain the FIRST occupied
            // sub-voxel on the ray's way (what an iterated "test per occupied sub-voxel" scheme would fetch first)
            Trav s;
            s.tmx = PF(F_TMX); s.tmy = PF(F_TMY); s.tmz = PF(F_TMZ); s.tdx = PF(F_TDX); s.tdy = PF(F_TDY); s.tdz = PF(F_TDZ);
            s.t0 = PF(F_T0); s.pos = PU(F_POS); s.flip = PU(F_FLIP); s.blin = (int32_t)PU(F_BLIN);
            s.wlo = s.whi = 0u; s.sx = s.sy = s.sz = 0;
            const uint4 rec = GSB_LDG_REC(g.cell_rec + trav_cell(s));
            Fine f;
            fine_enter(s, g, dx, dy, dz, f);
            unsigned long long path = 0ull, first = 0ull;
            const unsigned long long occ = ((unsigned long long)rec.w << 32) | rec.z;
            for (int guard = 0; guard < 16; ++guard) {
              const uint32_t bit = (((f.b * 0x1041u) >> 12) & 63u) ^ s.flip;
              path |= 1ull << bit;
              if (!first && ((occ >> bit) & 1ull)) first = 1ull << bit;
              const float t1 = fminf(f.fy, f.fz);
              const bool ax = f.fx <= t1, ay = !ax && f.fy <= f.fz, az = !ax && !ay;
              if (ax) { f.fx += f.fdx; f.b += 1u; }
              if (ay) { f.fy += f.fdy; f.b += 1u << 8; }
              if (az) { f.fz += f.fdz; f.b += 1u << 16; }
              if (f.b & 0x040404u) break;
            }
            const int nb4 = 4 * g.nb;
            const uint32_t local = trav_local(s);
            const int bx = s.blin % g.nb, by = (s.blin / g.nb) % g.nb, bz = s.blin / (g.nb * g.nb);
            const float lx = g.ox + (float)(4 * bx + (int)(local & 3u)) * g.cell, ly = g.oy + (float)(4 * by + (int)((local >> 2) & 3u)) * g.cell,
                        lz = g.oz + (float)(4 * bz + (int)((local >> 4) & 3u)) * g.cell;
            (void)nb4;
            const uint32_t n_here = min((uint32_t)kBatch, k1 - k0);
            for (uint32_t q = 0; q < n_here; ++q) {
              const float3 a = make_float3(ra[q].x, ra[q].y, ra[q].z);
              const float3 b = make_float3(a.x + ra[q].w, a.y + rb[q].x, a.z + rb[q].y);
              const float3 c = make_float3(a.x + rb[q].z, a.y + rb[q].w, a.z + rc[q]);
              const unsigned long long m = subvoxel_mask(a, b, c, lx, ly, lz, g.cell);
              if (m & path) GSB_STAT(6, 1);
              if (m & first) GSB_STAT(7, 1);
            }
          }
#endif
          k0 += kBatch;
          if (hit) {
            vis[PU(F_RID)] = 0;
            stt[slot] = (uint8_t)ST_IDLE;
            busy = false;
            GSB_STAT(3, 1);
          } else if (k0 >= k1) {
            stt[slot] = (uint8_t)ST_SEARCH;
            busy = false;
          }
        }
        if (GSB_TRACE_POOL_TROUNDS > 1 && __ballot_sync(full, busy) == 0u) break;
      }
      if (busy) PU(F_K0) = k0;
    }
  }
#undef PU
#undef PF
}

// Host copies of the grid descriptions built in this process (keyed by the device buffer): the trace kernel takes the struct
// by value.  Filled by gsb_occluder_build_fill, which already runs after the build's one host read.
struct OccCacheEntry { const void* dev; OccGrid g; };
OccCacheEntry g_occ_cache[16];
int g_occ_cache_next = 0;
void occ_cache_put(const void* dev, const OccGrid& g) {
  for (auto& e : g_occ_cache)
    if (e.dev == dev) { e.g = g; return; }
  g_occ_cache[g_occ_cache_next] = OccCacheEntry{dev, g};
  g_occ_cache_next = (g_occ_cache_next + 1) % 16;
}
bool occ_cache_get(const void* dev, OccGrid& g) {
  for (auto& e : g_occ_cache)
    if (e.dev == dev && dev != nullptr) { g = e.g; return true; }
  return false;
}

}  // namespace

extern "C" {

size_t gsb_occluder_struct_bytes(void) { return sizeof(OccGrid); }

int64_t gsb_occluder_brick_words(int grid_res) { const int64_t nb = (grid_res + 3) / 4; return nb * nb * nb; }

int64_t gsb_occluder_cells(int grid_res) { return 64 * gsb_occluder_brick_words(grid_res); }

int64_t gsb_occluder_scan_ws_ints(int64_t n_cells) { return (n_cells + kScanTile - 1) / kScanTile + 1; }

int gsb_occluder_build_count(const float* verts, const int32_t* tris, int64_t n_faces, const float* bounds_lo,
                             const float* bounds_hi, int grid_res, void* occluder, int32_t* cell_start, int32_t* scan_ws,
                             uint64_t* brick_bits, int32_t* total, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (grid_res < 1 || grid_res > kMaxGridRes) return (int)cudaErrorInvalidValue;      // 9-bit cell coordinates (trace_core.cuh)
  const int64_t n_cells = gsb_occluder_cells(grid_res);
  cudaError_t e = cudaMemsetAsync(cell_start, 0, sizeof(int32_t) * (size_t)(n_cells + 1), stream);
  if (e != cudaSuccess) return (int)e;
  k_params<<<1, 32, 0, stream>>>(bounds_lo, bounds_hi, grid_res, (OccGrid*)occluder, (const unsigned long long*)brick_bits);
  if (n_faces > 0)
    k_bin<false><<<nblk(n_faces), kThreads, 0, stream>>>(verts, tris, n_faces, (const OccGrid*)occluder, cell_start, nullptr, nullptr);
  k_brick_bits<<<nblk(n_cells), kThreads, 0, stream>>>(cell_start, n_cells, (uint32_t*)brick_bits);
  const int n_tiles = (int)((n_cells + kScanTile - 1) / kScanTile);
  k_scan_tiles<<<n_tiles, kThreads, 0, stream>>>(cell_start, n_cells, scan_ws);
  k_scan_sums<<<1, 1024, 0, stream>>>(scan_ws, n_tiles, total);
  k_scan_add<<<n_tiles, kThreads, 0, stream>>>(cell_start, n_cells, scan_ws, total);
  return (int)cudaGetLastError();
}

int gsb_occluder_build_fill(const float* verts, const int32_t* tris, int64_t n_faces, int grid_res, void* occluder,
                            const int32_t* cell_start, int32_t* cursor, void* cell_recs, float* cell_tri_data, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t n_cells = gsb_occluder_cells(grid_res);
  cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)n_cells, stream);
  if (e != cudaSuccess) return (int)e;
  k_cell_recs<<<nblk(n_cells), kThreads, 0, stream>>>(cell_start, n_cells, (uint4*)cell_recs);
  k_set_tables<<<1, 32, 0, stream>>>((OccGrid*)occluder, (const uint4*)cell_recs, (const float4*)cell_tri_data);
  if (n_faces > 0)
    k_bin<true><<<nblk(n_faces), kThreads, 0, stream>>>(verts, tris, n_faces, (const OccGrid*)occluder, cursor,
                                                        (float4*)cell_tri_data, (uint4*)cell_recs);
  if (n_faces > 0)
    k_subvoxels<<<148 * 8, kThreads, 0, stream>>>(verts, tris, (const OccGrid*)occluder, cell_start + n_cells,
                                                  (const float4*)cell_tri_data, (uint4*)cell_recs);
  // host copy of the (now complete) description for the trace launches; the stream was just synchronised by the caller's
  // read of *total, so this 56-byte copy waits only for k_cell_recs / k_set_tables
  OccGrid h;
  e = cudaMemcpyAsync(&h, occluder, sizeof(OccGrid), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return (int)e;
  occ_cache_put(occluder, h);
  return (int)cudaGetLastError();
}

int gsb_trace_shadow_rays(const void* occluder, const void* ray_list, const int32_t* ray_count, int64_t ray_cap,
                          int32_t* fetch_counter, uint8_t* vis, void* stream_) {
  OccGrid g;
  if (!occ_cache_get(occluder, g)) {      // built by another process / copied buffer: fetch the description once (synchronises)
    cudaError_t e = cudaMemcpy(&g, occluder, sizeof(OccGrid), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    occ_cache_put(occluder, g);
  }
  static bool pool_attr_set = false;
  if (!pool_attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_trace_pool, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSmemBytes);
    if (e != cudaSuccess) return (int)e;
    pool_attr_set = true;
  }
  // persistent grid: GSB_TRACE_POOL_BLOCKS CTAs per SM
  k_trace_pool<<<148 * GSB_TRACE_POOL_BLOCKS, GSB_TRACE_POOL_THREADS, kPoolSmemBytes, (cudaStream_t)stream_>>>(
      g, (const float4*)ray_list, ray_count, (int)(ray_cap < 0x70000000 ? ray_cap : 0x70000000), fetch_counter, vis);    // (head room: warps over-fetch the cursor by a block each)
  return (int)cudaGetLastError();
}

/* Traversal counters {triangle tests, cell steps, cells descended into, hits, sub-voxel steps, cells tested, -, -,
 * then for the pooled kernel: executions and summed active lanes of the SEARCH, DESC, TEST and refill blocks}; all zero
 * unless the library was built with -DGSB_TRACE_STATS (profiling builds only: the counters are global atomics). */
void gsb_trace_stats(uint64_t* out16, int reset) {
#ifdef GSB_TRACE_STATS
  unsigned long long v[16];
  cudaMemcpyFromSymbol(v, g_trace_stats, sizeof(v));
  for (int i = 0; i < 16; ++i) out16[i] = v[i];
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_trace_stats, z, sizeof(z)); }
#else
  (void)reset;
  for (int i = 0; i < 16; ++i) out16[i] = 0;
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
