// Shadow-ray traversal core shared by the trace kernel (occluder.cu) and the host-side logic test
// (tests/trace_host.cu): a three-level bit hierarchy walked with ONE lean 3-D DDA step routine.
//
//   level 0  bricks      4x4x4 cells      one 64-bit occupancy word per brick (1.3 MB at R = 217: cache-resident)
//   level 1  cells       the unit that owns a triangle list; one 16-byte record per cell, stored brick-major:
//                        {first entry, entries, 64-bit occupancy of the cell's 4x4x4 sub-voxels}
//   level 2  sub-voxels  bits only (no lists): a ray that crosses an occupied cell without touching an occupied
//                        sub-voxel never fetches a triangle
//
// Replaces optixTrace / the any-hit program of the reference (render/optixutils/c_src/envsampling/kernel.cu:101-118);
// B200 has no RT cores.  Why this shape (round-1 ncu, profiles/r1j): the old single-level walk spent 65 SASS
// instructions per cell step, tested 41 triangles per ray of which 91 % of the entered cells held no hit, and pulled
// 45x the algorithmic bytes from DRAM (duplicated 48-byte records fetched for every false-positive cell).
//
// The DDA runs in a MIRRORED frame: every direction component is made positive by reflecting the axis, so a step is
// always +1 / +4 / +16 on a 6-bit local index, "left the 4x4x4 block" is "the 2-bit field was 3", and the true bit
// index is `local ^ flip`.  Level 0/1 (trav_step) keeps 12 registers of state and touches memory only when a brick is
// left; level 2 (trav_descend) is a short walk on temporaries inside one cell.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GSB_HD __host__ __device__ __forceinline__
#else
#define GSB_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define GSB_LDG(p) __ldg(p)
#define GSB_RCP(x) __fdividef(1.f, (x))
#define GSB_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#define GSB_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
// 16-byte read-only load that does not allocate a line in L1 (profiling knob: keeps streamed records from evicting the brick words)
__device__ __forceinline__ uint4 gsb_ldg_na(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 gsb_ldg_na(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float gsb_ldg_na(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
#else
#define GSB_LDG(p) (*(p))
#define GSB_RCP(x) (1.f / (x))
#define GSB_PREFETCH_L1(p) ((void)0)
#define GSB_PREFETCH_L2(p) ((void)0)
#endif
// build-time latency knobs of the trace kernel (profiles/r2c: the walk is latency-bound, 41 % of the stall samples sit on the
// first use of a brick word / cell record / triangle record)
#ifndef GSB_TRACE_PF_BRICK
#define GSB_TRACE_PF_BRICK 0     // on entering a brick: prefetch the three bricks the ray can leave it into
#endif
#ifndef GSB_TRACE_PF_CELL
#define GSB_TRACE_PF_CELL 0      // on finding an occupied cell: prefetch its 16-byte record
#endif
#ifndef GSB_TRACE_NA_REC
#define GSB_TRACE_NA_REC 0       // 1: cell records bypass L1
#endif
#ifndef GSB_TRACE_NA_TRI
#define GSB_TRACE_NA_TRI 0       // 1: triangle records bypass L1
#endif
#if defined(__CUDA_ARCH__) && GSB_TRACE_NA_REC
#define GSB_LDG_REC(p) gsb_ldg_na(p)
#else
#define GSB_LDG_REC(p) GSB_LDG(p)
#endif
#if defined(__CUDA_ARCH__) && GSB_TRACE_NA_TRI
#define GSB_LDG_TRI(p) gsb_ldg_na(p)
#else
#define GSB_LDG_TRI(p) GSB_LDG(p)
#endif
#ifndef GSB_TRACE_PF_REC
#define GSB_TRACE_PF_REC 0       // on finding an occupied sub-voxel: prefetch the first triangle records of the cell (1 = L2, 2 = L1)
#endif

namespace gsb {

struct OccGrid {
  const unsigned long long* brick_occ;   // [nb^3] bit = (z&3)<<4 | (y&3)<<2 | (x&3), brick index (bz*nb + by)*nb + bx
  const uint4* cell_rec;                 // [nb^3 * 64] brick-major: {first entry, entries, sub-voxel mask lo, hi}
  const float4* tri_rec;                 // [entries][3] = (v0, e1, e2) of every (cell, triangle) pair, grouped by cell
  float ox, oy, oz;                      // grid origin (min corner)
  float cell, inv_cell;
  int n;                                 // cells per axis that can hold geometry (R)
  int nb;                                // bricks per axis = ceil(R / 4); cells past R exist as empty cells
};

// cell id in the brick-major order of cell_rec (and of the count / scan arrays of the build)
GSB_HD int64_t cell_id(int x, int y, int z, int nb) {
  return (((((int64_t)(z >> 2)) * nb + (y >> 2)) * nb + (x >> 2)) << 6) | (int64_t)(((z & 3) << 4) | ((y & 3) << 2) | (x & 3));
}

// ---- exact triangle / axis-aligned box overlap (separating axes; Akenine-Moeller) -----------------------------------
// vertices relative to the box centre, h = half extent
GSB_HD bool axis_separates(float ax, float ay, float az, float3 a, float3 b, float3 c, float h) {
  const float p0 = ax * a.x + ay * a.y + az * a.z, p1 = ax * b.x + ay * b.y + az * b.z, p2 = ax * c.x + ay * c.y + az * c.z;
  const float r = h * (fabsf(ax) + fabsf(ay) + fabsf(az));
  return fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r;
}
GSB_HD bool tri_overlaps_box(float3 a, float3 b, float3 c, float h) {
  const float3 e0 = make_float3(b.x - a.x, b.y - a.y, b.z - a.z), e1 = make_float3(c.x - b.x, c.y - b.y, c.z - b.z),
               e2 = make_float3(a.x - c.x, a.y - c.y, a.z - c.z);
  if (axis_separates(1.f, 0.f, 0.f, a, b, c, h) || axis_separates(0.f, 1.f, 0.f, a, b, c, h) || axis_separates(0.f, 0.f, 1.f, a, b, c, h)) return false;
  if (axis_separates(e0.y * e1.z - e0.z * e1.y, e0.z * e1.x - e0.x * e1.z, e0.x * e1.y - e0.y * e1.x, a, b, c, h)) return false;
  const float3 e[3] = {e0, e1, e2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (axis_separates(0.f, -e[k].z, e[k].y, a, b, c, h)) return false;
    if (axis_separates(e[k].z, 0.f, -e[k].x, a, b, c, h)) return false;
    if (axis_separates(-e[k].y, e[k].x, 0.f, a, b, c, h)) return false;
  }
  return true;
}

// Sub-voxel occupancy of triangle (a, b, c) inside the cell whose min corner is (lx, ly, lz): bit (sz<<4 | sy<<2 | sx) is set
// when the triangle overlaps the sub-voxel box grown by `kSubPad` of its half extent (the DDA's arithmetic places a point
// to ~1e-4 sub-voxels; the pad keeps the bits conservative against that).
constexpr float kSubPad = 1.02f;
// the 16 bits (sy<<2 | sx) of z-slice sz
GSB_HD uint32_t subvoxel_slice(float3 a, float3 b, float3 c, float lx, float ly, float lz, float cell, int sz) {
  const float sub = 0.25f * cell, inv = 4.f / cell, pad = (kSubPad - 1.f) * 0.5f * sub;
  int r0[3], r1[3];
  const float lo[3] = {fminf(a.x, fminf(b.x, c.x)) - lx, fminf(a.y, fminf(b.y, c.y)) - ly, fminf(a.z, fminf(b.z, c.z)) - lz};
  const float hi[3] = {fmaxf(a.x, fmaxf(b.x, c.x)) - lx, fmaxf(a.y, fmaxf(b.y, c.y)) - ly, fmaxf(a.z, fmaxf(b.z, c.z)) - lz};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r0[k] = (int)fminf(fmaxf(floorf((lo[k] - pad) * inv), 0.f), 3.f);
    r1[k] = (int)fminf(fmaxf(floorf((hi[k] + pad) * inv), 0.f), 3.f);
  }
  uint32_t m = 0u;
  if (sz < r0[2] || sz > r1[2]) return m;
  for (int y = r0[1]; y <= r1[1]; ++y)
    for (int x = r0[0]; x <= r1[0]; ++x) {
      const float cx = lx + (x + 0.5f) * sub, cy = ly + (y + 0.5f) * sub, cz = lz + (sz + 0.5f) * sub;
      if (tri_overlaps_box(make_float3(a.x - cx, a.y - cy, a.z - cz), make_float3(b.x - cx, b.y - cy, b.z - cz),
                           make_float3(c.x - cx, c.y - cy, c.z - cz), 0.5f * sub * kSubPad))
        m |= 1u << ((y << 2) | x);
    }
  return m;
}
GSB_HD unsigned long long subvoxel_mask(float3 a, float3 b, float3 c, float lx, float ly, float lz, float cell) {
  unsigned long long m = 0ull;
  for (int z = 0; z < 4; ++z) m |= (unsigned long long)subvoxel_slice(a, b, c, lx, ly, lz, cell, z) << (16 * z);
  return m;
}

// ---- ray / triangle: Moeller-Trumbore, two-sided, hit iff 0 < t < 1e16 (OptiX tmin = 0, tmax = 1e16; any-hit) ---------
// Division-free: u, v, t are compared after multiplying through by |det| (the quotient form 1/det costs a reciprocal with
// its slow path per test: 60 SASS instructions against ~38).  record (48 B): [v0.xyz e1.x] [e1.yz e2.xy] [e2.z]
GSB_HD bool ray_hits_triangle(const float4 ra, const float4 rb, const float rc, float ox, float oy, float oz, float dx, float dy,
                              float dz) {
  const float e1x = ra.w, e1y = rb.x, e1z = rb.y, e2x = rb.z, e2y = rb.w, e2z = rc;
  const float px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
  const float det = e1x * px + e1y * py + e1z * pz;
  const float tx = ox - ra.x, ty = oy - ra.y, tz = oz - ra.z;
  const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
  const float sg = det < 0.f ? -1.f : 1.f, ad = fabsf(det);
  const float u = (tx * px + ty * py + tz * pz) * sg;
  const float v = (dx * qx + dy * qy + dz * qz) * sg;
  const float t = (e2x * qx + e2y * qy + e2z * qz) * sg;
  return (ad > 0.f) & (u >= 0.f) & (v >= 0.f) & (u + v <= ad) & (t > 0.f) & (t < 1e16f * ad);
}

// ---- traversal state (levels 0/1) ---------------------------------------------------------------------------------------
// The cell position is ONE packed word in the mirrored frame, 10 bits per axis: bits 0-8 = coordinate + (512 - 4 nb), bit 9 =
// "past the last cell" guard (the bias makes the coordinate overflow into it exactly when the ray leaves the grid).  A step adds
// 1 / 1<<10 / 1<<20; "left the brick" = bit 2 of the stepped field flipped.  (Round-2 ncu, profiles/r2j: the kernel is bound by
// the half-rate ALU pipe -- selects, logic, compares -- so the step is written as three predicated add pairs, which issue on
// the FMA pipe, instead of select chains: 53 -> ~35 instructions per step.)
struct Trav {
  float tmx, tmy, tmz;      // time at which the ray leaves the current cell, per axis
  float tdx, tdy, tdz;      // time to cross one cell, per axis (finite: |d| is clamped away from 0)
  float t0;                 // time at which the ray enters the grid (>= 0): entry time of the FIRST cell only
  uint32_t pos;             // packed mirrored cell position (see above)
  uint32_t wlo, whi;        // occupancy word of the current brick
  uint32_t flip;            // 0b11 in the 2-bit field of every mirrored axis (6-bit brick-local layout z<<4 | y<<2 | x)
  int32_t blin;             // linear index of the current brick (true, un-mirrored)
  int32_t sx, sy, sz;       // signed brick strides of the three axes (derived from flip: trav_strides)
};
enum { TR_CONT = 0, TR_FOUND = 1, TR_EXIT = 2 };
constexpr float kMinDir = 1e-18f;
constexpr int kMaxGridRes = 512;                 // 9-bit coordinates
constexpr uint32_t kPosLocal = 0x00300C03u;      // low two bits of every field: the cell inside its brick
constexpr uint32_t kPosBit2 = 0x00401004u;       // bit 2 of every field: flips when a step wraps the brick-local coordinate
constexpr uint32_t kPosGuard = 0x20080200u;      // bit 9 of every field

GSB_HD bool word_bit(uint32_t lo, uint32_t hi, uint32_t i) { return ((((i & 32u) ? hi : lo) >> (i & 31u)) & 1u) != 0u; }
// brick-local index (true frame) of the current cell: the three 2-bit fields at bits 0, 10, 20 are gathered by one multiply
// (partial products land on disjoint bits; bits 16-21 of the product are z|y|x)
GSB_HD uint32_t trav_local(const Trav& s) { return ((((s.pos & kPosLocal) * 0x00010101u) >> 16) & 63u) ^ s.flip; }
GSB_HD bool trav_bit(const Trav& s) { return word_bit(s.wlo, s.whi, trav_local(s)); }
GSB_HD void trav_load_brick(Trav& s, const OccGrid& g) {
  const unsigned long long w = GSB_LDG(g.brick_occ + s.blin);
  s.wlo = (uint32_t)w;
  s.whi = (uint32_t)(w >> 32);
}
GSB_HD void trav_strides(Trav& s, const OccGrid& g) {
  const int nb = g.nb, nb2 = nb * nb;
  s.sx = (s.flip & 1u) ? -1 : 1;
  s.sy = (s.flip & 4u) ? -nb : nb;
  s.sz = (s.flip & 16u) ? -nb2 : nb2;
}
// entry time of the current cell: the largest of the three "previous boundary" times, and never before the grid entry
GSB_HD float trav_tcur(const Trav& s) { return fmaxf(fmaxf(s.tmx - s.tdx, s.tmy - s.tdy), fmaxf(s.tmz - s.tdz, s.t0)); }

// Clips the ray to the (brick-padded) grid box and sets up the walk in the cell of the entry point.
// Returns false when the ray misses the grid.  The caller then tests trav_bit() for the first cell.
GSB_HD bool trav_setup(Trav& s, const OccGrid& g, float ox, float oy, float oz, float dx, float dy, float dz) {
  const float cdx = fabsf(dx) < kMinDir ? kMinDir : dx, cdy = fabsf(dy) < kMinDir ? kMinDir : dy, cdz = fabsf(dz) < kMinDir ? kMinDir : dz;
  const float ix = GSB_RCP(cdx), iy = GSB_RCP(cdy), iz = GSB_RCP(cdz);
  const int nc = 4 * g.nb;
  const float ext = (float)nc * g.cell;
  float t0 = 0.f, t1 = 3.0e38f;
  {
    float a = (g.ox - ox) * ix, b = (g.ox + ext - ox) * ix;
    t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
    a = (g.oy - oy) * iy; b = (g.oy + ext - oy) * iy;
    t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
    a = (g.oz - oz) * iz; b = (g.oz + ext - oz) * iz;
    t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
  }
  if (!(t0 <= t1)) return false;
  const int cx = min(max((int)floorf((ox + dx * t0 - g.ox) * g.inv_cell), 0), nc - 1);
  const int cy = min(max((int)floorf((oy + dy * t0 - g.oy) * g.inv_cell), 0), nc - 1);
  const int cz = min(max((int)floorf((oz + dz * t0 - g.oz) * g.inv_cell), 0), nc - 1);
  const bool fx = cdx < 0.f, fy = cdy < 0.f, fz = cdz < 0.f;
  s.tmx = (g.ox + (float)(cx + (fx ? 0 : 1)) * g.cell - ox) * ix;
  s.tmy = (g.oy + (float)(cy + (fy ? 0 : 1)) * g.cell - oy) * iy;
  s.tmz = (g.oz + (float)(cz + (fz ? 0 : 1)) * g.cell - oz) * iz;
  s.tdx = g.cell * fabsf(ix);
  s.tdy = g.cell * fabsf(iy);
  s.tdz = g.cell * fabsf(iz);
  const int bias = kMaxGridRes - nc;
  const int mx = (fx ? nc - 1 - cx : cx) + bias, my = (fy ? nc - 1 - cy : cy) + bias, mz = (fz ? nc - 1 - cz : cz) + bias;
  s.pos = (uint32_t)mx | ((uint32_t)my << 10) | ((uint32_t)mz << 20);
  s.blin = ((cz >> 2) * g.nb + (cy >> 2)) * g.nb + (cx >> 2);
  s.flip = (fx ? 3u : 0u) | (fy ? 12u : 0u) | (fz ? 48u : 0u);
  s.t0 = t0;
  trav_strides(s, g);
  trav_load_brick(s, g);
  return true;
}

// One cell step.  TR_FOUND: the cell just entered is occupied.  TR_EXIT: the ray left the grid.
GSB_HD int trav_step(Trav& s, const OccGrid& g) {
  const float t1 = fminf(s.tmy, s.tmz);
  const bool ax = s.tmx <= t1;
  const bool ay = !ax && s.tmy <= s.tmz;
  const bool az = !ax && !ay;
  const uint32_t old = s.pos;
  if (ax) { s.tmx += s.tdx; s.pos += 1u; }
  if (ay) { s.tmy += s.tdy; s.pos += 1u << 10; }
  if (az) { s.tmz += s.tdz; s.pos += 1u << 20; }
  if ((s.pos ^ old) & kPosBit2) {               // left the brick
    if (s.pos & kPosGuard) return TR_EXIT;
    if (ax) s.blin += s.sx;
    if (ay) s.blin += s.sy;
    if (az) s.blin += s.sz;
    trav_load_brick(s, g);
#if GSB_TRACE_PF_BRICK
    const int last = g.nb * g.nb * g.nb - 1, nb2 = g.nb * g.nb;
    GSB_PREFETCH_L1(g.brick_occ + min(max(s.blin + ((s.flip & 1u) ? -1 : 1), 0), last));
    GSB_PREFETCH_L1(g.brick_occ + min(max(s.blin + ((s.flip & 4u) ? -g.nb : g.nb), 0), last));
    GSB_PREFETCH_L1(g.brick_occ + min(max(s.blin + ((s.flip & 16u) ? -nb2 : nb2), 0), last));
#endif
  }
  const bool found = trav_bit(s);
#if GSB_TRACE_PF_CELL
  if (found) GSB_PREFETCH_L1(g.cell_rec + (((int64_t)s.blin << 6) | (int64_t)trav_local(s)));
#endif
  return found ? TR_FOUND : TR_CONT;
}

// Level 2: the walk stands in an occupied cell.  The cell's 4x4x4 sub-voxel bits are walked from the entry point to the cell's
// exit on temporaries (at most 10 sub-voxels).  FINE_HIT at the first occupied sub-voxel -- the caller then tests the cell's
// triangles -- FINE_EXIT when the ray leaves the cell without touching one (no triangle is fetched), FINE_MORE when `max_steps`
// ran out first (the caller keeps `b` and resumes: the trace kernel bounds the loop because it runs until the slowest lane of
// the warp is through).  The level-1 state is not modified.
enum { FINE_EXIT = 0, FINE_HIT = 1, FINE_MORE = 2 };
struct Fine {
  float fx, fy, fz;         // time at which the ray leaves the current sub-voxel, per axis
  float fdx, fdy, fdz;      // time to cross one sub-voxel
  uint32_t b;               // sub-voxel position: 2-bit fields at bits 0 / 8 / 16, a guard bit above each (set when a step leaves the cell)
};
GSB_HD void fine_place(const Trav& s, float qx, float qy, float qz, Fine& f) {      // q = sub-voxels still ahead on each axis (0..3)
  f.fdx = 0.25f * s.tdx; f.fdy = 0.25f * s.tdy; f.fdz = 0.25f * s.tdz;
  f.fx = fmaf(-qx, f.fdx, s.tmx); f.fy = fmaf(-qy, f.fdy, s.tmy); f.fz = fmaf(-qz, f.fdz, s.tmz);
}
GSB_HD void fine_enter(const Trav& s, const OccGrid& g, float dx, float dy, float dz, Fine& f) {
  const float tcur = trav_tcur(s);
  // fraction of the cell still ahead of the entry point, in sub-voxels (mirrored frame: the ray moves towards +)
  const float k = 4.f * g.inv_cell;
  const float qx = fminf(fmaxf(floorf((s.tmx - tcur) * (fmaxf(fabsf(dx), kMinDir) * k)), 0.f), 3.f);
  const float qy = fminf(fmaxf(floorf((s.tmy - tcur) * (fmaxf(fabsf(dy), kMinDir) * k)), 0.f), 3.f);
  const float qz = fminf(fmaxf(floorf((s.tmz - tcur) * (fmaxf(fabsf(dz), kMinDir) * k)), 0.f), 3.f);
  fine_place(s, qx, qy, qz, f);
  f.b = (uint32_t)(3 - (int)qx) | ((uint32_t)(3 - (int)qy) << 8) | ((uint32_t)(3 - (int)qz) << 16);
}
GSB_HD void fine_resume(const Trav& s, uint32_t b, Fine& f) {
  fine_place(s, (float)(3u - (b & 3u)), (float)(3u - ((b >> 8) & 3u)), (float)(3u - ((b >> 16) & 3u)), f);
  f.b = b;
}
GSB_HD int fine_walk(uint32_t mlo, uint32_t mhi, uint32_t flip, Fine& f, int max_steps, uint32_t& steps) {
  steps = 0;
  for (;;) {
    // the 6-bit index z<<4 | y<<2 | x comes out of one multiply (partial products on disjoint bits, bits 12-17 of the product)
    if (word_bit(mlo, mhi, (((f.b * 0x1041u) >> 12) & 63u) ^ flip)) return FINE_HIT;
    if ((int)steps >= max_steps) return FINE_MORE;
    const float t1 = fminf(f.fy, f.fz);
    const bool ax = f.fx <= t1;
    const bool ay = !ax && f.fy <= f.fz;
    const bool az = !ax && !ay;
    if (ax) { f.fx += f.fdx; f.b += 1u; }
    if (ay) { f.fy += f.fdy; f.b += 1u << 8; }
    if (az) { f.fz += f.fdz; f.b += 1u << 16; }
    if (f.b & 0x040404u) return FINE_EXIT;       // the step left the cell
    ++steps;
  }
}
GSB_HD int64_t trav_cell(const Trav& s) { return ((int64_t)s.blin << 6) | (int64_t)trav_local(s); }

// unbounded descent (host driver, single-pass callers): true at an occupied sub-voxel
GSB_HD bool trav_descend(const Trav& s, const OccGrid& g, float dx, float dy, float dz, uint32_t& first, uint32_t& count,
                         uint32_t& fine_steps) {
  const uint4 rec = GSB_LDG_REC(g.cell_rec + trav_cell(s));
  first = rec.x;
  count = rec.y;
  Fine f;
  fine_enter(s, g, dx, dy, dz, f);
  return fine_walk(rec.z, rec.w, s.flip, f, 1 << 30, fine_steps) == FINE_HIT;
}

}  // namespace gsb
