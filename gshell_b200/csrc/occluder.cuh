// Uniform-grid occluder for shadow (any-hit) rays against the just-extracted mesh.
//
// The reference rebuilds an OptiX GAS every iteration (render/optixutils/c_src/torch_bindings.cpp:37-116) and
// traces one terminate-on-first-hit ray per light/BSDF sample (envsampling/kernel.cu:101-118).  B200 has no RT
// cores and this path may not use OptiX, so the acceleration structure is designed for THIS geometry: the mesh is
// a soup of 10^5..10^7 near-uniformly sized micro-triangles emitted from a regular tet grid, for which a uniform
// grid with cells about one tet wide is near-optimal and builds with count -> scan -> fill (no sort, no tree).
// Traversal = 3-D DDA (Amanatides & Woo) with early exit on the first hit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsb {

struct Occluder {
  const int32_t* cell_start;   // [nx*ny*nz + 1] exclusive prefix of per-cell triangle counts
  const int32_t* cell_tris;    // triangle ids, grouped by cell
  const float4* tri_data;      // [F][3] = (v0, e1 = v1-v0, e2 = v2-v0), w unused
  const unsigned long long* brick_mask;   // [(nx/4)*(ny/4)*(nz/4)] bit (x&3)+4(y&3)+16(z&3) = cell holds triangles
  float ox, oy, oz;            // grid origin (min corner)
  float inv_cell;              // 1 / cell size
  float cell;                  // cell size
  int nx, ny, nz;
};

// Moeller-Trumbore, two-sided, hit iff t > 0 (OptiX tmin = 0, tmax = 1e16; any-hit)
__device__ __forceinline__ bool ray_hits_triangle(const float4* __restrict__ td, float ox, float oy, float oz, float dx,
                                                  float dy, float dz) {
  const float4 v0 = __ldg(td), e1 = __ldg(td + 1), e2 = __ldg(td + 2);
  const float px = dy * e2.z - dz * e2.y, py = dz * e2.x - dx * e2.z, pz = dx * e2.y - dy * e2.x;
  const float det = e1.x * px + e1.y * py + e1.z * pz;
  if (det == 0.f) return false;
  const float inv = 1.f / det;
  const float tx = ox - v0.x, ty = oy - v0.y, tz = oz - v0.z;
  const float u = (tx * px + ty * py + tz * pz) * inv;
  if (u < 0.f || u > 1.f) return false;
  const float qx = ty * e1.z - tz * e1.y, qy = tz * e1.x - tx * e1.z, qz = tx * e1.y - ty * e1.x;
  const float v = (dx * qx + dy * qy + dz * qz) * inv;
  if (v < 0.f || u + v > 1.f) return false;
  const float t = (e2.x * qx + e2.y * qy + e2.z * qz) * inv;
  return t > 0.f && t < 1e16f;
}

// true if any triangle blocks the ray (o, d), d need not be normalised.
// Two-level 3-D DDA: bricks of 4x4x4 cells carry a 64-bit occupancy mask, empty bricks are crossed in one step and
// empty cells inside a brick cost one bit test (no memory access).
__device__ __forceinline__ bool occluded(const Occluder& g, float ox, float oy, float oz, float dx, float dy, float dz) {
  const float bx = g.ox + g.nx * g.cell, by = g.oy + g.ny * g.cell, bz = g.oz + g.nz * g.cell;
  const float idx = 1.f / dx, idy = 1.f / dy, idz = 1.f / dz;      // +-inf for axis-parallel rays is handled below
  float t0 = 0.f, t1 = 3.0e38f;
  const bool inside = ox >= g.ox && ox <= bx && oy >= g.oy && oy <= by && oz >= g.oz && oz <= bz;
  if (!inside) {     // surface points start inside the grid: the slab clip is the rare path
    float a = (g.ox - ox) * idx, b = (bx - ox) * idx;
    if (dx == 0.f) { if (ox < g.ox || ox > bx) return false; } else { t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    a = (g.oy - oy) * idy; b = (by - oy) * idy;
    if (dy == 0.f) { if (oy < g.oy || oy > by) return false; } else { t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    a = (g.oz - oz) * idz; b = (bz - oz) * idz;
    if (dz == 0.f) { if (oz < g.oz || oz > bz) return false; } else { t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    if (!(t0 <= t1)) return false;
  }
  const int sx = dx > 0.f ? 1 : -1, sy = dy > 0.f ? 1 : -1, sz = dz > 0.f ? 1 : -1;
  const int px = sx > 0 ? 1 : 0, py = sy > 0 ? 1 : 0, pz = sz > 0 ? 1 : 0;
  const float big = 3.0e38f, bsz = 4.f * g.cell, inv_b = 0.25f * g.inv_cell;
  const int nbx = g.nx >> 2, nby = g.ny >> 2, nbz = g.nz >> 2;
  // entry brick
  float ex = ox + dx * t0, ey = oy + dy * t0, ez = oz + dz * t0;
  int Bx = min(max((int)floorf((ex - g.ox) * inv_b), 0), nbx - 1);
  int By = min(max((int)floorf((ey - g.oy) * inv_b), 0), nby - 1);
  int Bz = min(max((int)floorf((ez - g.oz) * inv_b), 0), nbz - 1);
  float TX = dx != 0.f ? (g.ox + (Bx + px) * bsz - ox) * idx : big;
  float TY = dy != 0.f ? (g.oy + (By + py) * bsz - oy) * idy : big;
  float TZ = dz != 0.f ? (g.oz + (Bz + pz) * bsz - oz) * idz : big;
  const float DX = dx != 0.f ? bsz * fabsf(idx) : big, DY = dy != 0.f ? bsz * fabsf(idy) : big, DZ = dz != 0.f ? bsz * fabsf(idz) : big;
  const float tdx = 0.25f * DX, tdy = 0.25f * DY, tdz = 0.25f * DZ;
  float t_enter = t0;
  for (;;) {
    const unsigned long long mask = __ldg(g.brick_mask + ((size_t)Bz * nby + By) * nbx + Bx);
    if (mask) {
      // fine DDA through the cells of this brick, starting where the ray enters it
      ex = ox + dx * t_enter; ey = oy + dy * t_enter; ez = oz + dz * t_enter;
      int cx = min(max((int)floorf((ex - g.ox) * g.inv_cell), 4 * Bx), 4 * Bx + 3);
      int cy = min(max((int)floorf((ey - g.oy) * g.inv_cell), 4 * By), 4 * By + 3);
      int cz = min(max((int)floorf((ez - g.oz) * g.inv_cell), 4 * Bz), 4 * Bz + 3);
      float tmx = dx != 0.f ? (g.ox + (cx + px) * g.cell - ox) * idx : big;
      float tmy = dy != 0.f ? (g.oy + (cy + py) * g.cell - oy) * idy : big;
      float tmz = dz != 0.f ? (g.oz + (cz + pz) * g.cell - oz) * idz : big;
      for (;;) {
        const int bit = (cx & 3) | ((cy & 3) << 2) | ((cz & 3) << 4);
        if ((mask >> bit) & 1ull) {
          const int c = (cz * g.ny + cy) * g.nx + cx;
          const int b0 = __ldg(g.cell_start + c), b1 = __ldg(g.cell_start + c + 1);
          for (int k = b0; k < b1; ++k) {
            const int f = __ldg(g.cell_tris + k);
            if (ray_hits_triangle(g.tri_data + (size_t)f * 3, ox, oy, oz, dx, dy, dz)) return true;
          }
        }
        if (tmx <= tmy && tmx <= tmz) { cx += sx; if ((cx >> 2) != Bx) break; tmx += tdx; }
        else if (tmy <= tmz)          { cy += sy; if ((cy >> 2) != By) break; tmy += tdy; }
        else                          { cz += sz; if ((cz >> 2) != Bz) break; tmz += tdz; }
      }
    }
    if (TX <= TY && TX <= TZ) { t_enter = TX; Bx += sx; if (Bx < 0 || Bx >= nbx) return false; TX += DX; }
    else if (TY <= TZ)        { t_enter = TY; By += sy; if (By < 0 || By >= nby) return false; TY += DY; }
    else                      { t_enter = TZ; Bz += sz; if (Bz < 0 || Bz >= nbz) return false; TZ += DZ; }
  }
}

}  // namespace gsb
