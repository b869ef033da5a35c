// Uniform-grid occluder for shadow (any-hit) rays against the just-extracted mesh.
//
// The reference rebuilds an OptiX GAS every iteration (render/optixutils/c_src/torch_bindings.cpp:37-116) and
// traces one terminate-on-first-hit ray per light/BSDF sample (envsampling/kernel.cu:101-118).  B200 has no RT
// cores and this path may not use OptiX, so the acceleration structure is designed for THIS geometry: the mesh is
// a soup of 10^5..10^7 near-uniformly sized micro-triangles emitted from a regular tet grid, for which a uniform
// grid with cells about one tet wide is near-optimal and builds with count -> scan -> fill (no sort, no tree).
// Traversal = 3-D DDA (Amanatides & Woo) with early exit on the first hit (k_trace_list in occluder.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsb {

struct Occluder {
  const int32_t* cell_start;   // [nx*ny*nz + 1] exclusive prefix of per-cell triangle counts
  const float4* cell_tri_data; // [entries][3] = (v0, e1 = v1-v0, e2 = v2-v0) of every (cell, triangle) pair, grouped by cell:
                               // triangle data is duplicated per cell so a cell visit is two dependent loads, not three
  const unsigned long long* brick_occ;  // [nbz][nby][nbx] occupancy bits of 4x4x4-cell bricks, bit = (z&3)<<4 | (y&3)<<2 | (x&3):
                               // 1/32 of the range table, L2-resident, so stepping through an EMPTY cell never touches HBM
  int nbx, nby;                // bricks per axis (ceil(n/4))
  const uint32_t* cell_slabs;  // [cells] bits 0-7 / 8-15 / 16-23: eighth-of-a-cell slabs along x / y / z touched by the cell's
                               // triangles (clipped AABBs) => a sub-cell bounding box; a ray that misses it skips the whole cell
  float ox, oy, oz;            // grid origin (min corner)
  float inv_cell;              // 1 / cell size
  float cell;                  // cell size
  int nx, ny, nz;
};

// Moeller-Trumbore, two-sided, hit iff t > 0 (OptiX tmin = 0, tmax = 1e16; any-hit);
// branch-free variant for SIMD-uniform inner loops: every lane executes the same ~40 instructions
// record layout (48 B): [v0.x v0.y v0.z e1.x] [e1.y e1.z e2.x e2.y] [e2.z - - -]: two 16-byte loads and one 4-byte load
__device__ __forceinline__ bool ray_hits_triangle_bf(const float4 ra, const float4 rb, const float rc, float ox, float oy,
                                                     float oz, float dx, float dy, float dz) {
  const float3 v0 = make_float3(ra.x, ra.y, ra.z), e1 = make_float3(ra.w, rb.x, rb.y), e2 = make_float3(rb.z, rb.w, rc);
  const float px = dy * e2.z - dz * e2.y, py = dz * e2.x - dx * e2.z, pz = dx * e2.y - dy * e2.x;
  const float det = e1.x * px + e1.y * py + e1.z * pz;
  const float inv = 1.f / det;
  const float tx = ox - v0.x, ty = oy - v0.y, tz = oz - v0.z;
  const float u = (tx * px + ty * py + tz * pz) * inv;
  const float qx = ty * e1.z - tz * e1.y, qy = tz * e1.x - tx * e1.z, qz = tx * e1.y - ty * e1.x;
  const float v = (dx * qx + dy * qy + dz * qz) * inv;
  const float t = (e2.x * qx + e2.y * qy + e2.z * qz) * inv;
  return (det != 0.f) & (u >= 0.f) & (u <= 1.f) & (v >= 0.f) & (u + v <= 1.f) & (t > 0.f) & (t < 1e16f);
}

}  // namespace gsb
