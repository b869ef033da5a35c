// Host-side driver of the shadow-ray traversal core (gshell_b200/csrc/trace_core.cuh) for the CPU test suite: builds the
// three-level occluder of a triangle soup with the same functions the CUDA build kernels call, then walks rays one at a
// time through the same SEARCH / DESC / TEST state machine as k_trace_pool (occluder.cu).  Test infrastructure only (tests/test_trace_host.py);
// what it cannot cover is the warp-level glue of the kernel (ballots, refill, atomics), which the -m gpu tests exercise.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../gshell_b200/csrc/trace_core.cuh"

using namespace gsb;

namespace {
struct HostOcc {
  OccGrid g;
  std::vector<unsigned long long> bricks;
  std::vector<uint4> recs;
  std::vector<float4> tris;
};

float3 ldv(const float* v, int i) { return make_float3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); }

void build(HostOcc& h, const float* verts, int64_t n_verts, const int32_t* tris, int64_t F, int R) {
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int64_t i = 0; i < n_verts; ++i)
    for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], verts[3 * i + k]); hi[k] = fmaxf(hi[k], verts[3 * i + k]); }
  OccGrid& o = h.g;
  const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), fmaxf(hi[2] - lo[2], 1e-6f));
  o.cell = ext * 1.0001f / (float)R;
  o.inv_cell = 1.f / o.cell;
  o.ox = 0.5f * (lo[0] + hi[0]) - 0.5f * o.cell * R;
  o.oy = 0.5f * (lo[1] + hi[1]) - 0.5f * o.cell * R;
  o.oz = 0.5f * (lo[2] + hi[2]) - 0.5f * o.cell * R;
  o.n = R;
  o.nb = (R + 3) / 4;
  const int64_t n_cells = (int64_t)o.nb * o.nb * o.nb * 64;
  std::vector<int> count(n_cells + 1, 0);
  struct Entry { int64_t cell; int64_t tri; };
  std::vector<Entry> entries;
  const float pad = 1e-4f * o.cell;
  for (int64_t f = 0; f < F; ++f) {
    const float3 a = ldv(verts, tris[3 * f]), b = ldv(verts, tris[3 * f + 1]), c = ldv(verts, tris[3 * f + 2]);
    const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
    if (uy * vz - uz * vy == 0.f && uz * vx - ux * vz == 0.f && ux * vy - uy * vx == 0.f) continue;
    int r0[3], r1[3];
    const float mn[3] = {fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z))};
    const float mx[3] = {fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z))};
    const float og[3] = {o.ox, o.oy, o.oz};
    for (int k = 0; k < 3; ++k) {
      r0[k] = std::min(std::max((int)floorf((mn[k] - pad - og[k]) * o.inv_cell), 0), R - 1);
      r1[k] = std::min(std::max((int)floorf((mx[k] + pad - og[k]) * o.inv_cell), 0), R - 1);
    }
    for (int z = r0[2]; z <= r1[2]; ++z)
      for (int y = r0[1]; y <= r1[1]; ++y)
        for (int x = r0[0]; x <= r1[0]; ++x) {
          const float ccx = o.ox + (x + 0.5f) * o.cell, ccy = o.oy + (y + 0.5f) * o.cell, ccz = o.oz + (z + 0.5f) * o.cell;
          if (!tri_overlaps_box(make_float3(a.x - ccx, a.y - ccy, a.z - ccz), make_float3(b.x - ccx, b.y - ccy, b.z - ccz),
                                make_float3(c.x - ccx, c.y - ccy, c.z - ccz), 0.5f * o.cell * 1.001f))
            continue;
          const int64_t cid = cell_id(x, y, z, o.nb);
          ++count[cid];
          entries.push_back({cid, f});
        }
  }
  h.recs.assign(n_cells, make_uint4(0, 0, 0, 0));
  h.bricks.assign((size_t)o.nb * o.nb * o.nb, 0ull);
  int run = 0;
  for (int64_t i = 0; i < n_cells; ++i) {
    h.recs[i].x = run; h.recs[i].y = count[i];
    if (count[i] > 0) h.bricks[i >> 6] |= 1ull << (i & 63);
    run += count[i];
  }
  h.tris.assign((size_t)run * 3 + 3, make_float4(0, 0, 0, 0));
  std::vector<int> cursor(n_cells, 0);
  for (const Entry& e : entries) {
    const int64_t f = e.tri;
    const float3 a = ldv(verts, tris[3 * f]), b = ldv(verts, tris[3 * f + 1]), c = ldv(verts, tris[3 * f + 2]);
    // inverse of cell_id
    const int64_t br = e.cell >> 6;
    const int l = (int)(e.cell & 63);
    const int x = (int)(br % o.nb) * 4 + (l & 3), y = (int)((br / o.nb) % o.nb) * 4 + ((l >> 2) & 3), z = (int)(br / ((int64_t)o.nb * o.nb)) * 4 + (l >> 4);
    const unsigned long long m = subvoxel_mask(a, b, c, o.ox + x * o.cell, o.oy + y * o.cell, o.oz + z * o.cell, o.cell);
    h.recs[e.cell].z |= (uint32_t)m;
    h.recs[e.cell].w |= (uint32_t)(m >> 32);
    const size_t s = 3 * (size_t)(h.recs[e.cell].x + cursor[e.cell]++);
    h.tris[s] = make_float4(a.x, a.y, a.z, b.x - a.x);
    h.tris[s + 1] = make_float4(b.y - a.y, b.z - a.z, c.x - a.x, c.y - a.y);
    h.tris[s + 2] = make_float4(c.z - a.z, 0.f, 0.f, 0.f);
  }
  o.brick_occ = h.bricks.data();
  o.cell_rec = h.recs.data();
  o.tri_rec = h.tris.data();
}
}  // namespace

extern "C" {

// rays float[N,6] (origin, direction); vis uint8[N] = 1 visible / 0 occluded.
// stats int64[6]: entries, triangle tests, cell steps, sub-voxel steps, cells descended into, cells tested.
// use_subvoxels = 0: every occupied cell reached is tested (the level-2 bits are ignored): isolates the level-0/1 walk;
// 1: unbounded sub-voxel walk; n >= 2: walk cut into pieces of n - 1 steps and resumed.
int trace_host(const float* verts, int64_t n_verts, const int32_t* tris, int64_t n_faces, int grid_res, const float* rays,
               int64_t n_rays, int use_subvoxels, uint8_t* vis, int64_t* stats) {
  HostOcc h;
  build(h, verts, n_verts, tris, n_faces, grid_res);
  const OccGrid& g = h.g;
  memset(stats, 0, 6 * sizeof(int64_t));
  stats[0] = (int64_t)h.tris.size() / 3 - 1;
  for (int64_t j = 0; j < n_rays; ++j) {
    const float ox = rays[6 * j], oy = rays[6 * j + 1], oz = rays[6 * j + 2], dx = rays[6 * j + 3], dy = rays[6 * j + 4], dz = rays[6 * j + 5];
    vis[j] = 1;
    Trav s;
    memset(&s, 0, sizeof(s));
    if (!trav_setup(s, g, ox, oy, oz, dx, dy, dz)) continue;
    int st = trav_bit(s) ? 1 : 0;
    uint32_t k0 = 0, k1 = 0;
    for (int guard = 0; guard < 1000000; ++guard) {
      if (st == 2) {
        bool hit = false;
        for (; k0 < k1 && !hit; ++k0) {
          const float4* t = g.tri_rec + 3 * (size_t)k0;
          hit = ray_hits_triangle(t[0], t[1], t[2].x, ox, oy, oz, dx, dy, dz);
          ++stats[1];
        }
        if (hit) { vis[j] = 0; break; }
        st = 0;
      } else if (st == 1) {
        ++stats[4];
        // use_subvoxels >= 2: the walk is cut into pieces of (use_subvoxels - 1) steps and resumed from the saved position, as the
        // trace kernel does with GSB_TRACE_FINE_CAP; the result must not depend on the cut
        const uint4 rec = g.cell_rec[trav_cell(s)];
        const uint32_t first = rec.x, count = rec.y;
        const int cap = use_subvoxels > 1 ? use_subvoxels - 1 : (1 << 30);
        Fine f;
        fine_enter(s, g, dx, dy, dz, f);
        int r;
        for (;;) {
          uint32_t fs;
          r = fine_walk(rec.z, rec.w, s.flip, f, cap, fs);
          stats[3] += fs;
          if (r != FINE_MORE) break;
          const uint32_t saved = f.b;
          fine_resume(s, saved, f);
        }
        const bool occ = r == FINE_HIT;
        if (occ || !use_subvoxels) {
          k0 = first; k1 = first + count;
          st = 2;
          ++stats[5];
        } else {
          st = 0;
        }
      } else {
        ++stats[2];
        const int r = trav_step(s, g);
        if (r == TR_EXIT) break;
        if (r == TR_FOUND) st = 1;
      }
    }
  }
  return 0;
}

}  // extern "C"
