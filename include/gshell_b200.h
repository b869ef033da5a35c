/* gshell_b200 -- C ABI of the B200-native (sm_100a) G-Shell inverse-rendering hot path.
 *
 * Every entry point takes raw DEVICE pointers (unless a name ends in `_host`), plain sizes and a
 * `cudaStream_t` passed as `void*`; nothing here depends on torch.  All functions are asynchronous
 * on `stream` and return 0 on success or a `cudaError_t` value (as int) on failure.
 * The Python host layer (`gshell_b200/`) binds these through ctypes; INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   gsb_mt_*          geometry/gshell_tets.py:245-443      GShell_Tets.__call__
 *   gsb_xfm_points_*  render/renderutils/c_src/mesh.cu:22,56 via torch_bindings.cpp:971-1004
 *   ... (see each section)
 */
#ifndef GSHELL_B200_H
#define GSHELL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Library info
 * ---------------------------------------------------------------------------------------------- */
/* Returns the ABI version (bumped on any signature change). */
int gsb_abi_version(void);
/* Compiled SM architecture, e.g. 100 for sm_100a. */
int gsb_compiled_arch(void);

/* ------------------------------------------------------------------------------------------------
 * G-Shell marching tetrahedra (replaces GShell_Tets.__call__, geometry/gshell_tets.py:245-443)
 *
 * Static topology tables (built once per tet grid by the host layer, see
 * gshell_b200/geometry/tet_tables.py):
 *   tet_v  int32[T,4]  vertex ids of each tet (the reference's tet_fx4, narrowed to int32)
 *   tet_e  int32[T,6]  for the 6 tet edges in local order (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
 *                      (gshell_tets.py:178) the id of that edge in `edge_v`
 *   edge_v int32[E,2]  all unique grid edges as (lo,hi), sorted lexicographically -- the order
 *                      `torch.unique(all_edges, dim=0)` produces at gshell_tets.py:268
 *
 * Extraction runs in two phases around ONE host read of `counts` (needed to size the outputs):
 *   gsb_mt_count  -> fills `counts` (device, int32[GSB_MT_NCOUNTS]) and the workspace
 *   gsb_mt_emit   -> writes the exactly-sized outputs
 * ---------------------------------------------------------------------------------------------- */
enum {
  GSB_MT_VW = 0,      /* watertight vertices (= crossing edges)                       */
  GSB_MT_T1 = 1,      /* tets producing 1 watertight triangle (triangle polygons)     */
  GSB_MT_T2 = 2,      /* tets producing 2 watertight triangles (quad polygons)        */
  GSB_MT_G0 = 3,      /* polygons per cut group: tri->1, tri->2, quad->1..4 triangles */
  GSB_MT_NCOUNTS = 16
};

size_t gsb_mt_workspace_bytes(int64_t n_tets, int64_t n_edges);

int gsb_mt_count(const float* pos, const float* sdf, const float* msdf,   /* [Nv,3], [Nv], [Nv] */
                 const int32_t* tet_v, const int32_t* tet_e, const int32_t* edge_v,
                 int64_t n_verts, int64_t n_tets, int64_t n_edges,
                 void* workspace, size_t workspace_bytes,
                 int watertight_template,                           /* 1: reference default; 0: output_watertight_template=False
                                                                       (gshell_tets.py:260-263): tets without a positive mSDF corner are
                                                                       dropped before the edge numbering                              */
                 int32_t* counts,                                   /* device int32[16]      */
                 void* stream);

/* Outputs (sizes from counts: Vw, T1, T2, G0..G5):
 *   verts_aug  float[Va,3], Va = Vw + 3*T1 + 4*T2   (unreferenced rows are zero, :419-423)
 *   msdf_aug   float[Va]    extra['msdf'] (:386-390)
 *   faces_aug  int32[Fa,3], Fa = G0 + 2*G1 + G2 + 2*G3 + 3*G4 + 4*G5, six groups in order (:409-416)
 *   verts_wt   float[Vw,3]  extra['vertices_watertight']
 *   faces_wt   int32[Fw,3], Fw = T1 + 2*T2 (:313-316)
 *   vert_edge  int32[Vw]    edge id (row of edge_v) of each watertight vertex   (saved for backward)
 *   slot_a     int32[Va-Vw] first polygon vertex of each boundary slot; bit 31 = row is referenced
 */
int gsb_mt_emit(const float* pos, const float* sdf,                /* [Nv,3], [Nv]          */
                const int32_t* tet_e, const int32_t* edge_v,
                int64_t n_tets, int64_t n_edges,
                const void* workspace, const int32_t* counts,
                float* verts_aug, float* msdf_aug, int32_t* faces_aug,
                float* verts_wt, int32_t* faces_wt, int32_t* vert_edge, int32_t* slot_a,
                void* stream);

/* Analytic backward.  Any of g_verts_aug / g_msdf_aug / g_verts_wt may be NULL (treated as zero).
 * g_pos [Nv,3], g_sdf [Nv], g_msdf [Nv] must be zero-initialised by the caller; scratch is
 * float[Vw,5], zero-initialised by this call.  Stop-gradient structure follows
 * gshell_tets.py:290,383-384. */
int gsb_mt_backward(const float* pos, const float* sdf, const float* msdf,
                    const int32_t* edge_v,
                    const float* verts_wt, const float* msdf_aug,
                    const int32_t* vert_edge, const int32_t* slot_a,
                    int64_t n_verts_wt, int64_t n_tri_polys, int64_t n_quad_polys,
                    const float* g_verts_aug, const float* g_msdf_aug, const float* g_verts_wt,
                    float* scratch,
                    float* g_pos, float* g_sdf, float* g_msdf,
                    void* stream);


/* ------------------------------------------------------------------------------------------------
 * G-buffer operators of render/renderutils (reference: c_src/torch_bindings.cpp:971-1004 xfm_fwd/bwd,
 * :33-120 prepare_shading_normal_fwd/bwd, :800-900 image_loss_fwd/bwd; kernels mesh.cu, normal.cu, loss.cu)
 * ---------------------------------------------------------------------------------------------- */
/* out[b,n,:] = matrix[b] * (points[n],1).  points [1,N,3] (points_batched=0) or [B,N,3]; out [B,N,4]. */
int gsb_xfm_points_fwd(const float* points, const float* matrix, int64_t n_batch, int64_t n_points,
                       int points_batched, float* out, void* stream);
int gsb_xfm_points_bwd(const float* matrix, const float* g_out, int64_t n_batch, int64_t n_points,
                       int points_batched, float* g_points, void* stream);

/* inputs: HOST array of 6 device pointers (pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm);
 * strides: HOST array [6][3] = (batch, row, pixel) strides in floats, 0 for broadcast dims.
 * out / g_out / g_inputs[k] are dense [B,H,W,3]; g_inputs[k] may be NULL (gradient not needed). */
int gsb_shading_normal_fwd(const float* const* inputs, const int64_t* strides, int64_t B, int64_t H, int64_t W,
                           int two_sided, int opengl, float* out, void* stream);
int gsb_shading_normal_bwd(const float* const* inputs, const int64_t* strides, int64_t B, int64_t H, int64_t W,
                           int two_sided, int opengl, const float* g_out, float* const* g_inputs, void* stream);

/* loss: 0 l1, 1 mse, 2 relmse, 3 smape; tonemapper: 0 none, 1 log_srgb.  img/target dense, n_values = B*H*W*3.
 * fwd writes gsb_image_loss_partials(n_values) block sums (of the channel-mean loss); the scalar loss is
 * sum(partials) / (B*H*W).  bwd: g = *g_scalar * scale * dloss/dvalue (scale = 1/(B*H*W)). */
int64_t gsb_image_loss_partials(int64_t n_values);
int gsb_image_loss_fwd(const float* img, const float* target, int64_t n_values, int loss, int tonemapper,
                       float* partials, void* stream);
int gsb_image_loss_bwd(const float* img, const float* target, int64_t n_values, int loss, int tonemapper,
                       const float* g_scalar, float scale, float* g_img, float* g_target, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Monte-Carlo environment shading (reference: render/optixutils/c_src/torch_bindings.cpp:123-272
 * env_shade_fwd/bwd -> OptiX raygen envsampling/kernel.cu:463).  All tensors dense fp32:
 *   mask [B,H,W]; ro,pos,nrm,kd,ks [B,H,W,3]; view_pos [B,3]; light [lh,lw,3]; pdf,cols [lh,lw]; rows [lh];
 *   perms int32 [n_perms, n^2].  bsdf: 0 pbr, 1 diffuse, 2 white.  bvh: NULL = no occluders.
 *   Shadow rays (bvh != NULL and shadow_scale > 0) are traced wavefront-style: the sample directions of a chunk of
 *   sample pairs are regenerated into `scratch`, traced by gsb_trace_shadow_rays at full SIMD occupancy, and consumed by the
 *   shading pass; scratch_bytes = gsb_env_shade_scratch_bytes(B,H,W,n,budget) (>= 16 sample pairs, ideally all n^2).
 *   vis_bits: optional uint32 [B*H*W, ceil(2 n^2 / 32)]: fwd records the shadow-ray result of every sample (bit =
 *   visible), bwd replays it instead of tracing again (valid when fwd and bwd use the same rnd_seed); NULL = trace.
 *   rows_top [16], cols_top [lh,16]: optional (NULL = binary search) every-16th-entry tables of the CDFs, padded with
 *   2.0, enabling a 16-ary search with two 64-byte loads (needs lh, lw multiples of 16 and <= 256).
 * bwd zero-initialises g_light itself -- float[lh,lw,4], RGB in the first three channels of 16-byte texels so that every sample
 * adds its contribution with ONE vector reduction (red.global.add.v4.f32); g_pos,g_nrm,g_kd,g_ks are fully written.
 * ---------------------------------------------------------------------------------------------- */
/* Profiling aid: 1 = start recording CUDA events around the shadow-trace launches of the following env_shade calls;
 * 0 = stop, synchronise and return their summed device time in ms. */
float gsb_trace_timing(int enable);
/* Trace launches since the last gsb_trace_timing(1); the summed time covers the first 1024 of them. */
int gsb_trace_launches(void);
/* Rays handed to the trace kernel since the last reset (profiling aid; synchronises the device). */
uint64_t gsb_trace_ray_count(int reset);
/* 16 traversal counters {triangle tests, cell steps, cells descended into, hits, sub-voxel steps, cells tested, -, -, then
 * executions and summed active lanes of the SEARCH / DESC / TEST / refill blocks of the pooled kernel}: zeros unless built
 * with -DGSB_TRACE_STATS. */
void gsb_trace_stats(uint64_t* out16, int reset);
/* n_covered = an upper bound of the pixels with mask > 0 (0 or >= B*H*W: all pixels): the ray list of a chunk is sized for
 * 2 rays per covered pixel and sample pair, so sparse views need fewer, larger chunks.  A ray past the capacity (n_covered
 * understated) is counted; from then on every traced gsb_env_shade_* call returns cudaErrorInvalidValue until
 * gsb_env_shade_dropped_rays(1) acknowledges it (the call that lost the rays has already returned: its outputs are invalid). */
uint32_t gsb_env_shade_dropped_rays(int reset);
size_t gsb_env_shade_scratch_bytes(int64_t B, int64_t H, int64_t W, int64_t n_covered, int n_samples_x, size_t budget_bytes);
int gsb_env_shade_chunks(int64_t B, int64_t H, int64_t W, int64_t n_covered, int n_samples_x, size_t scratch_bytes);
/* Any-hit trace of a compact ray list (2 float4 per ray: (origin, ray id as int bits), (direction, -)); min(*ray_count,
 * ray_cap) rays; fetch_counter: device int zeroed by the caller; vis uint8[...] pre-set to 1, vis[ray id] = 0 on a hit. */
int gsb_trace_shadow_rays(const void* occluder, const void* ray_list, const int32_t* ray_count, int64_t ray_cap,
                          int32_t* fetch_counter, uint8_t* vis, void* stream);
/* pixel_ids (may be NULL): uint32[B*H*W], the number that seeds pixel i's sample stream instead of i itself.  The forward and
 * the backward pass of one pixel must see the same (rnd_seed, id); a caller that shades pixels gathered from several images --
 * the row-block exchange of the multi-GPU step, render/optixutils/ops.py -- passes the ids the pixels have at home. */
int gsb_env_shade_fwd(const float* mask, const float* ro, const float* pos, const float* nrm, const float* view_pos,
                      const float* kd, const float* ks, const float* light, const float* pdf, const float* rows,
                      const float* cols, const float* rows_top, const float* cols_top, const int32_t* perms, int64_t B, int64_t H,
                      int64_t W, int64_t lh, int64_t lw, int64_t n_perms, int bsdf, int n_samples_x, uint32_t rnd_seed,
                      float shadow_scale, const void* bvh, void* scratch, size_t scratch_bytes, int64_t n_covered, uint32_t* vis_bits, float* diff,
                      float* spec, const uint32_t* pixel_ids, void* stream);
int gsb_env_shade_bwd(const float* mask, const float* ro, const float* pos, const float* nrm, const float* view_pos,
                      const float* kd, const float* ks, const float* light, const float* pdf, const float* rows,
                      const float* cols, const float* rows_top, const float* cols_top, const int32_t* perms, int64_t B, int64_t H,
                      int64_t W, int64_t lh, int64_t lw, int64_t n_perms, int bsdf, int n_samples_x, uint32_t rnd_seed,
                      float shadow_scale, const void* bvh, void* scratch, size_t scratch_bytes, int64_t n_covered, const uint32_t* vis_bits,
                      const float* g_diff, const float* g_spec, float* g_pos, float* g_nrm,
                      float* g_kd, float* g_ks, float* g_light, const uint32_t* pixel_ids, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multiresolution hash-grid encoding of the learned material field (reference render/mlptexture.py:49-83, where it is
 * tcnn.Encoding(3, {"otype": "HashGrid", ...}) of tiny-cuda-nn -- third-party and absent: own implementation of the
 * published algorithm, parity unpinned).  x01 float[n,3] in [0,1]; table float[level_offset[n_levels]][2] (all levels,
 * 2 features per entry); level_offset uint32[n_levels+1], level_res uint32[n_levels], level_scale float[n_levels] are HOST
 * arrays (n_levels <= 32); out / g_out float[n, 2*n_levels].  bwd zeroes and fills g_table (same shape as table) and
 * writes g_x float[n,3]; either may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int gsb_hashgrid_fwd(const float* x01, int64_t n, const float* table, const uint32_t* level_offset, const uint32_t* level_res,
                     const float* level_scale, int n_levels, float* out, void* stream);
int gsb_hashgrid_bwd(const float* x01, int64_t n, const float* table, const uint32_t* level_offset, const uint32_t* level_res,
                     const float* level_scale, int n_levels, const float* g_out, float* g_table, float* g_x, void* stream);
/* The whole material field at inference time, one launch (MLPTexture3D.sample under torch.no_grad(), reference render/mlptexture.py:86-98):
 * clamp((pos - lo) / (hi - lo), 0, 1) -> hash-grid encoding (n_levels = 16, width 32) -> Linear(32,32) ReLU Linear(32,32) ReLU
 * Linear(32,n_out), all bias-free, weights row-major [out][in] as torch.nn.Linear.weight -> sigmoid(.) * (max - min) + min.
 * aabb_lo_hi6_host = HOST {lo.xyz, hi.xyz}; out_lo_hi_host = HOST {min[n_out], max[n_out]}; n_out <= 16; out float[n, n_out]. */
int gsb_field_infer(const float* pos, int64_t n, const float* table, const uint32_t* level_offset, const uint32_t* level_res,
                    const float* level_scale, int n_levels, const float* w1, const float* w2, const float* w3, int n_out,
                    const float* aabb_lo_hi6_host, const float* out_lo_hi_host, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cross-bilateral denoiser (reference: render/optixutils/c_src/denoising.cu:14,74 via
 * torch_bindings.cpp:274-318).  col_* / nrm / zdz are [B,H,W,C] views given by base pointer + pixel
 * stride in floats (rows and batches dense); col_b may be NULL, otherwise both images are filtered in
 * one pass with shared weights.  Outputs: out_* = sum(w c)/max(sum w,1e-4) [B,H,W,3], w_* [B,H,W].
 * bwd: g_col = adjoint gather of g_out / w; inv_w_scratch is float[2*B*H*W].
 * ---------------------------------------------------------------------------------------------- */
int gsb_bilateral_fwd(const float* col_a, const float* col_b, const float* nrm, const float* zdz, int64_t ps_col,
                      int64_t ps_nrm, int64_t ps_zdz, int64_t B, int64_t H, int64_t W, float sigma, float* out_a,
                      float* w_a, float* out_b, float* w_b, void* stream);
int gsb_bilateral_bwd(const float* g_out_a, const float* w_a, const float* g_out_b, const float* w_b, const float* nrm,
                      const float* zdz, int64_t ps_nrm, int64_t ps_zdz, int64_t B, int64_t H, int64_t W, float sigma,
                      float* g_col_a, float* g_col_b, float* inv_w_scratch, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Rasteriser + attribute interpolation (replaces the nvdiffrast calls of the reference:
 * render/render.py:377-379 DepthPeeler.rasterize_next_layer (first layer), :26 dr.interpolate).
 * Conventions (nvdiffrast's): clip [1|B,V,4]; tris int32 [F,3]; rast [B,H,W,4] = (u, v, z/w, id+1) with u,v the
 * perspective-correct barycentrics of vertices 0 and 1; image row j is NDC y = (j+0.5)/H*2-1; rast_db
 * [B,H,W,4] = (du/dX, du/dY, dv/dX, dv/dY) per pixel (may be NULL).  zbuf: scratch uint64[B*H*W].
 * Triangles with any w <= 1e-8 are culled (no clipper); depth outside [-1,1] is clipped per pixel.
 * rasterize_bwd accumulates (atomics) d(u,v) into g_clip [1|B,V,4] (x,y,w components), caller zero-inits.
 * interpolate: attr [1|B,V,C] -> out [B,H,W,C] (0 where empty); out_d [B,H,W,C,2] optional (needs rast_db).
 * interpolate_bwd: g_attr (caller zero-inits, may be NULL) += ..., g_rast [B,H,W,4] written (may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int gsb_rasterize_fwd(const float* clip, const int32_t* tris, int64_t n_batch, int64_t n_verts, int64_t n_tris,
                      int clip_batched, int64_t H, int64_t W, void* zbuf, float* rast, float* rast_db, void* stream);
int gsb_rasterize_bwd(const float* clip, const int32_t* tris, const float* rast, const float* g_rast, int64_t n_batch,
                      int64_t n_verts, int clip_batched, int64_t H, int64_t W, float* g_clip, void* stream);
int gsb_interpolate_fwd(const float* attr, const int32_t* tris, const float* rast, const float* rast_db, int64_t n_batch,
                        int64_t n_verts, int64_t n_channels, int attr_batched, int64_t H, int64_t W, float* out,
                        float* out_d, void* stream);
int gsb_interpolate_bwd(const float* attr, const int32_t* tris, const float* rast, const float* g_out, int64_t n_batch,
                        int64_t n_verts, int64_t n_channels, int attr_batched, int64_t H, int64_t W, float* g_attr,
                        float* g_rast, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Smooth vertex normals (replaces mesh.auto_normals, reference render/mesh.py:212-237).
 * acc float[V,3] = unnormalised area-weighted sums (kept for backward); normals float[V,3].
 * bwd: g_acc float[V,3] scratch, g_verts float[V,3] (zero-initialised by the call).
 * ---------------------------------------------------------------------------------------------- */
int gsb_vertex_normals_fwd(const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_faces, float* acc,
                           float* normals, void* stream);
int gsb_vertex_normals_bwd(const float* verts, const int32_t* tris, const float* acc, const float* g_normals,
                           int64_t n_verts, int64_t n_faces, float* g_acc, float* g_verts, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Tangent frame of the extracted mesh (replaces map_uv / compute_tangents and the extension to the boundary vertices,
 * reference geometry/gshell_tets.py:40-78, 210-239, 318-319, 337-338, 375-380).
 *   verts float[Vw,3], faces int32[Fw,3], normals float[Vw,3] (gsb_vertex_normals_fwd): the watertight mesh
 *   msdf_wt float[Vw]   interpolated mSDF of the watertight vertices; slot_a int32[nb]: per boundary vertex the watertight vertex
 *                       its polygon edge starts at (bit 31 ignored), 3 per triangle polygon first (n_tri_polys of them), then
 *                       4 per quad polygon, as gsb_mt_emit writes them; both may be NULL when n_boundary = 0
 *   atlas_lin float[atlas_n] = linspace(0, 1 - 1/atlas_n, atlas_n), atlas_n = ceil(sqrt(T)), atlas_pad = 0.9 / atlas_n: the uv
 *                       atlas of map_uv is arithmetic on the row number, no table is built
 *   acc float[Vw,4]     per-vertex tangent sum + face count (kept for backward); tng_aug float[Vw+nb,3]
 * bwd: g_tng_aug float[Vw+nb,3] in; g_t float[Vw,3] scratch; g_verts float[Vw,3], g_normals float[Vw,3], g_msdf float[Vw] out
 * (fully written by the call).
 * ---------------------------------------------------------------------------------------------- */
int gsb_tangents_fwd(const float* verts, const int32_t* faces, const float* normals, const float* msdf_wt, const int32_t* slot_a,
                     const float* atlas_lin, int64_t n_wt, int64_t n_faces, int64_t n_tri_polys, int64_t n_boundary,
                     int32_t atlas_n, float atlas_pad, float* acc, float* tng_aug, void* stream);
int gsb_tangents_bwd(const int32_t* faces, const float* normals, const float* msdf_wt, const int32_t* slot_a, const float* atlas_lin,
                     int64_t n_wt, int64_t n_faces, int64_t n_tri_polys, int64_t n_boundary, int32_t atlas_n, float atlas_pad,
                     const float* acc, const float* tng_aug, const float* g_tng_aug, float* g_t, float* g_verts, float* g_normals,
                     float* g_msdf, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Decode of a generated augmented grid (replaces GShell_Tets.marching_from_auggrid, reference geometry/gshell_tets.py:446-629;
 * no gradients there either).  Static tables tet_v / tet_e / edge_v as for gsb_mt_*.  The host layer runs two int32 prefix sums
 * between the calls (flags -> flags_incl over the edges; rows -> rows_incl along the tets of each of the 8 rows) and reads two
 * sets of totals to size the outputs.  Pointers named *_host are HOST arrays.
 *   edge_flags : flags int32[E] = 1 where the SDF changes sign on the edge; vertex id of such an edge = flags_incl[e] - 1
 *   vertices   : verts float[Vw,3] (generated interpolation coefficient, coeff_grid float[gx,gy,gz] at the integer mid-point of the
 *                canonical edge, verts_discretized float[Nv,3]), verts_cano float[Vw,3] (canonical mid-point), msdf_vert float[Vw]
 *   classify   : rows int32[8,T]: {triangle polygon, quad polygon, cut group 0..5 = triangle -> 1,2 faces, quad -> 1,2,3,4 faces};
 *                luts7_host = device pointers to int32 {tri[16,6], loop[16,4], ntri[16], cut3[8,6], cut4[16,12], ncut3[8], ncut4[16]}
 *                (gshell_b200/geometry/mt_luts.py, negative entries clamped to 0)
 *   emit       : totals8_host = last column of rows_incl; faces_wt int32[n_one + 2 n_two, 3], tet_ids int32[n_one + n_two],
 *                boundary_pos float[nb,3], boundary_ab int32[nb,2] + boundary_w float[nb,2] (end points / weights of every
 *                boundary vertex, nb = 3 n_one + 4 n_two; occgrid float[ox,oy,oz] on the doubled grid), faces_aug int32[Fa,3] in
 *                the reference's six groups, Fa = sum over the groups of polygons x faces
 *   boundary_attr : out float[nb,3] = attr[a] w0 + attr[b] w1 (the tangents of the boundary vertices)
 * ---------------------------------------------------------------------------------------------- */
int gsb_auggrid_edge_flags(const float* sdf, const int32_t* edge_v, int64_t n_edges, int32_t* flags, void* stream);
int gsb_auggrid_vertices(const float* pos, const float* verts_discretized, const int32_t* edge_v, const int32_t* flags,
                         const int32_t* flags_incl, int64_t n_edges, const float* coeff_grid, const float* msdf_sign_grid,
                         int32_t gx, int32_t gy, int32_t gz, float* verts, float* verts_cano, float* msdf_vert, void* stream);
int gsb_auggrid_classify(const float* sdf, const int32_t* tet_v, const int32_t* tet_e, const int32_t* flags, const int32_t* flags_incl,
                         const float* msdf_vert, int64_t n_tets, const int32_t* const* luts7_host, int32_t* rows, void* stream);
int gsb_auggrid_emit(const float* sdf, const int32_t* tet_v, const int32_t* tet_e, const int32_t* flags, const int32_t* flags_incl,
                     const float* msdf_vert, const float* verts, const float* verts_cano, const int32_t* rows_incl, int64_t n_tets,
                     const int32_t* const* luts7_host, const float* occgrid, int32_t ox, int32_t oy, int32_t oz, int64_t n_wt,
                     const int64_t* totals8_host, int32_t* faces_wt, int32_t* tet_ids, float* boundary_pos, int32_t* boundary_ab,
                     float* boundary_w, int32_t* faces_aug, void* stream);
int gsb_auggrid_boundary_attr(const float* attr, const int32_t* boundary_ab, const float* boundary_w, int64_t n_boundary, float* out,
                              void* stream);


/* ------------------------------------------------------------------------------------------------
 * Occluder for shadow rays (replaces optix_build_bvh, reference render/optixutils/c_src/torch_bindings.cpp:37-116, and the
 * optixTrace any-hit query, envsampling/kernel.cu:101-118): a three-level bit hierarchy over the mesh bounds -- 4x4x4-cell
 * bricks (one 64-bit occupancy word each), grid_res^3 cells owning triangle lists, 4x4x4 sub-voxel bits per cell --
 * built count -> scan -> fill around one host read of *total (number of (cell, triangle) entries).  `occluder` is a device
 * buffer of gsb_occluder_struct_bytes() bytes; pass it as `bvh` to gsb_env_shade_*.  With n_cells = gsb_occluder_cells(grid_res)
 * (cells padded to whole bricks, stored brick-major): cell_start int32[n_cells+1]; scan_ws int32[gsb_occluder_scan_ws_ints(n_cells)];
 * brick_bits uint64[gsb_occluder_brick_words(grid_res)]: the only table an empty cell touches; cursor int32[n_cells];
 * cell_recs 16 bytes x n_cells {first entry, entries, sub-voxel bits}; cell_tri_data float[*total * 12] (triangle records
 * v0,e1,e2 duplicated per overlapped cell so that a cell's list is one contiguous read; words 9-10 of a record hold the cell and
 * triangle ids the build needs for the sub-voxel pass, word 11 is unused).  1 <= grid_res <= 512.
 * ---------------------------------------------------------------------------------------------- */
size_t gsb_occluder_struct_bytes(void);
int64_t gsb_occluder_cells(int grid_res);
int64_t gsb_occluder_scan_ws_ints(int64_t n_cells);
int64_t gsb_occluder_brick_words(int grid_res);
int gsb_occluder_build_count(const float* verts, const int32_t* tris, int64_t n_faces, const float* bounds_lo,
                             const float* bounds_hi, int grid_res, void* occluder, int32_t* cell_start, int32_t* scan_ws,
                             uint64_t* brick_bits, int32_t* total, void* stream);
int gsb_occluder_build_fill(const float* verts, const int32_t* tris, int64_t n_faces, int grid_res, void* occluder,
                            const int32_t* cell_start, int32_t* cursor, void* cell_recs, float* cell_tri_data, void* stream);


/* ------------------------------------------------------------------------------------------------
 * G-FlexiCubes topology (integer / ordering stages of GShellFlexiCubes.__call__, reference
 * geometry/gshell_flexicubes.py:136-230; see csrc/flexicubes.cu).  Static tables of the voxel grid (host layer,
 * gshell_b200/geometry/flex_tables.py): cube_v int32[C,8]; edge_v int32[E,2] sorted unique ORIENTED edges (:86-87,
 * :316-317); cube_e int32[C,12]; edge_cnt uint8[E] cubes around each edge; edge_slots int32[E,4] their
 * (cube*12+local edge) slots ascending.  LUTs (gshell_b200/geometry/flexicubes_tables.npz): check_table int16[256,5],
 * num_vd_table int8[256], dmc_table int8[256,4,7], gflex ntri int8[8], gflex configuration int8[8,6].
 * Phase 1 gsb_fc_count -> counts int32[8] = {surface cubes, cubes with 1,2,3,4 dual vertices, crossing edges,
 * flipped quads, regular quads}; phase 2 gsb_fc_emit writes surf_edges[n_cross,2], per-dual-vertex records
 * (vd_cube, vd_rank, vd_le int8[n_vd,7], vd_ce int32[n_vd,7] crossing-edge ids) and quads int32[Q,4].
 * gsb_fc_cut_count / _emit: open-surface cut of the triangulated quads (:554-591): counts int32[4] = {uncut, cut,
 * cut->1 tri, cut->2 tri}; faces_open int32[uncut + cut1 + 2 cut2, 3], cut_faces int32[cut,3].
 * blk_* are scratch: int32[5*gsb_fc_blocks(C)], int32[3*gsb_fc_blocks(E)], int32[4*gsb_fc_blocks(F)].
 * ---------------------------------------------------------------------------------------------- */
int64_t gsb_fc_blocks(int64_t n);
int gsb_fc_count(const float* s, const int32_t* cube_v, const int32_t* edge_v, const uint8_t* edge_cnt, const int16_t* check_table,
                 const int8_t* num_vd_table, int64_t n_cubes, int64_t n_edges, int res, uint8_t* raw_case, uint8_t* case_id,
                 int32_t* blk_cubes, int32_t* blk_edges, int32_t* counts, void* stream);
int gsb_fc_emit(const float* s, const int32_t* cube_e, const int32_t* edge_v, const uint8_t* edge_cnt, const int32_t* edge_slots,
                const int8_t* dmc_table, const int8_t* num_vd_table, int64_t n_cubes, int64_t n_edges, const uint8_t* case_id,
                const int32_t* blk_cubes, const int32_t* blk_edges, const int32_t* counts, int32_t* edge_cid, int32_t* quad_row,
                int32_t* slot_vd, int32_t* surf_edges, int32_t* vd_cube, int32_t* vd_rank, int8_t* vd_le, int32_t* vd_ce,
                int32_t* quads, int64_t n_flip, void* stream);
int gsb_fc_cut_count(const int32_t* faces, const float* nu_d, int64_t n_faces, const int8_t* ntri_table, int32_t* blk,
                     int32_t* counts, void* stream);
int gsb_fc_cut_emit(const int32_t* faces, const float* nu_d, int64_t n_faces, const int8_t* ntri_table, const int8_t* conf_table,
                    const int32_t* blk, const int32_t* counts, int64_t n_vd, int32_t* faces_open, int32_t* cut_faces,
                    void* stream);


/* FlexiCubes floating-point stages with analytic adjoints (csrc/flexicubes.cu):
 *   dual: per dual vertex, position vd, interpolated mSDF nu_d (with the reference's in-place aliasing, :476-477) and its
 *         stop-gradient twin, L_dev (:232-240) at l_off[v].. (l_off = exclusive prefix of the slot counts); alpha [C,8] and
 *         beta [C,12] are the NORMALISED weights (:242-263).  bwd accumulates (atomics) into zero-initialised g_* arrays.
 *   boundary: 3 mSDF zero-crossing vertices per cut face (:569-577), weights detached for the value path (:574). */
int gsb_fc_dual_fwd(const float* x, const float* s, const float* nu, const float* alpha, const float* beta,
                    const int32_t* surf_edges, const int32_t* vd_cube, const int8_t* vd_le, const int32_t* vd_ce,
                    const int32_t* l_off, int64_t n_vd, float* vd, float* nu_d, float* nu_d_sg, float* l_dev, void* stream);
int gsb_fc_dual_bwd(const float* x, const float* s, const float* nu, const float* alpha, const float* beta,
                    const int32_t* surf_edges, const int32_t* vd_cube, const int8_t* vd_le, const int32_t* vd_ce,
                    const int32_t* l_off, int64_t n_vd, const float* g_vd, const float* g_nu_d, const float* g_nu_d_sg,
                    const float* g_l_dev, float* g_x, float* g_s, float* g_nu, float* g_alpha, float* g_beta, void* stream);
int gsb_fc_boundary_fwd(const int32_t* cut_faces, int64_t n_cut, const float* vd, const float* nu_d, const float* nu_d_sg,
                        float* bverts, float* bnu_sg, void* stream);
int gsb_fc_boundary_bwd(const int32_t* cut_faces, int64_t n_cut, const float* vd, const float* nu_d, const float* nu_d_sg,
                        const float* g_bverts, const float* g_bnu_sg, float* g_vd, float* g_nu_d, float* g_nu_d_sg,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Light-probe tables and the loss assembly of one training iteration (csrc/tick_ops.cu).
 *
 * gsb_light_pdf: EnvironmentLight.update_pdf (reference render/light.py:46-59): pdf = max(rgb) sin(theta) normalised over the
 *   probe, cols = per-row column CDF (last entry 1), rows = row CDF replicated over the row (the reference's [h,w] layout).
 *   base float[h,w,3]; row_sum_ws float[h] scratch; pdf, cols, rows float[h,w].
 * gsb_sdf_reg_*: compute_sdf_reg_loss (geometry/gshell_tets_geometry.py:33-39) over a static edge table int32[E,2]:
 *   acc2 = {sum of both BCE terms over sign-changing edges, their count}; loss = acc2[0] / acc2[1].  bwd ADDS
 *   weight * *g_loss * d loss / d sdf into g_sdf (atomics).
 * gsb_mark_visible_boundary / gsb_msdf_reg_*: the two mSDF Huber regularisers (gshell_tets_geometry.py:325-356):
 *   acc2 = {sum huber(max(msdf, -eps) + eps) over all vertices, sum huber(min(msdf_b, eps) - eps) over the boundary vertices
 *   of visible triangles}; bmask uint8[n_boundary] is zeroed by the caller and set here from the visible-triangle id list.
 * gsb_image_terms_*: alpha MSE + the two mSDF-image L1 terms (gshell_tets_geometry.py:283-290), chroma_loss, shading_loss,
 *   material_smoothness_grad (render/regularizer.py:21-52) in one pass over the composited buffers ([n_pix,4] each, msdf_img
 *   [n_pix,msdf_ch]; NULL = term input absent).  terms: bit 0 alpha, 1 mSDF image, 2 chroma, 3 shading, 4 smoothness.
 *   lambdas6 (host) = {chroma, diffuse, specular, kd, ks, nrm}.  reduce -> acc double[gsb_image_terms_accumulators()]
 *   (the caller may all-reduce acc[5], acc[6] = sums of specular / diffuse luma over ranks and pass mean_scale = 1/world) ->
 *   finish -> out2 = {image part, regulariser part}.  bwd writes every non-NULL gradient buffer completely.
 * ---------------------------------------------------------------------------------------------- */
int gsb_light_pdf(const float* base, int64_t h, int64_t w, float* row_sum_ws, float* pdf, float* cols, float* rows, void* stream);
int gsb_sdf_reg_fwd(const float* sdf, const int32_t* edges, int64_t n_edges, double* acc2, void* stream);
int gsb_sdf_reg_bwd(const float* sdf, const int32_t* edges, int64_t n_edges, const double* acc2, const float* g_loss, float weight,
                    float* g_sdf, void* stream);
int gsb_mark_visible_boundary(const int32_t* tris, const int64_t* visible_ids, int64_t n_visible, int64_t n_verts_watertight,
                              uint8_t* bmask, void* stream);
int gsb_msdf_reg_fwd(const float* msdf_all, int64_t n_all, const float* msdf_boundary, const uint8_t* bmask, int64_t n_boundary, float eps,
                     double* acc2, void* stream);
int gsb_msdf_reg_bwd(const float* msdf_all, int64_t n_all, const float* msdf_boundary, const uint8_t* bmask, int64_t n_boundary, float eps,
                     const float* g_loss, float w_open, float w_close, float* g_all, float* g_boundary, void* stream);
int gsb_image_terms_accumulators(void);
int gsb_image_terms_reduce(const float* shaded, const float* msdf_img, const float* kd, const float* kd_grad, const float* ks_grad,
                           const float* nrm_grad, const float* diffuse, const float* specular, const float* ref, int64_t n_pix, int msdf_ch,
                           int terms, const float* lambdas6, double* acc, void* stream);
int gsb_image_terms_finish(const float* msdf_img, const float* diffuse, const float* specular, int64_t n_pix, int msdf_ch, int terms,
                           const float* lambdas6, const double* acc, float mean_scale, float* out2, void* stream);
int gsb_image_terms_bwd(const float* shaded, const float* msdf_img, const float* kd, const float* kd_grad, const float* ks_grad,
                        const float* nrm_grad, const float* diffuse, const float* specular, const float* ref, int64_t n_pix, int msdf_ch,
                        int terms, const float* lambdas6, const double* acc, float mean_scale, const float* g_out2, float* g_shaded,
                        float* g_msdf_img, float* g_kd, float* g_kd_grad, float* g_ks_grad, float* g_nrm_grad, float* g_diffuse,
                        float* g_specular, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused G-buffer build and fused shade-combine + composite (csrc/gbuffer_fused.cu).
 *
 * gsb_gbuffer_*: what render_layer assembles from five dr.interpolate calls and ~25 tensor ops (reference render/render.py:
 *   236-285): per pixel, from the rasteriser output rast/rast_db [B,H,W,4] and ONE gather of the covering triangle's vertices:
 *   interpolated position and vertex normal [B,H,W,3], unit face normal [B,H,W,3], depth (z/w and its one-pixel change)
 *   [B,H,W,2] (not differentiable), interpolated mSDF [B,H,W] (msdf NULL = skip).  v_clip float[B,V,4].
 *   bwd: g_v_pos, g_v_nrm, g_msdf are accumulated (zero them first), g_rast [B,H,W,4] = (d/du, d/dv, 0, 0) is written; NULL
 *   gradients in or out are skipped.
 * gsb_compose_*: shade()'s combine + buffer assembly + render_mesh()'s composite over the background (render.py:55-63,100-118,
 *   144-186,352-433) in one pass.  mode 0 'pbr' (diff kd (1-metal) + spec), 1 'diffuse' (diff kd), 2 colour given, 3 'white'.
 *   composite 1: every buffer is laid over the background (black for all but `shaded`) with the coverage as alpha, as
 *   render_mesh returns them; 0: alpha = 1 everywhere, as shade() / render_layer() return them.
 *   in12  = {rast, jitter[.,2]|NULL, gb_nrm[.,3], tex[.,6], tex_jittered[.,6], shading_nrm[.,3], geo_nrm[.,3], depth[.,2],
 *            diff[.,3]|NULL, spec[.,3]|NULL, col[.,3]|NULL, msdf_img[.]|NULL};  bg [B|1,H,W,3] or NULL (black).
 *   out12 = {shaded, z_grad, normal, geometric_normal, kd, ks, kd_grad, ks_grad, normal_grad, diffuse_light, specular_light}
 *           [B,H,W,4] each and msdf_image [B,H,W,1]; NULL entries are not produced.
 *   bwd: gin9 = gradients of {shaded, kd, ks, kd_grad, ks_grad, normal_grad, diffuse_light, specular_light, msdf_image};
 *   gout7 = {diff, spec, col, tex, tex_jittered, gb_nrm (accumulated: zero it first), msdf_img}; NULL entries are skipped.
 * ---------------------------------------------------------------------------------------------- */
int gsb_gbuffer_fwd(const float* rast, const float* rast_db, const float* v_pos, const float* v_nrm, const float* msdf, const float* v_clip,
                    const int32_t* tris, int64_t n_batch, int64_t H, int64_t W, int64_t n_verts, float* pos, float* nrm, float* geo_nrm,
                    float* depth, float* msdf_img, void* stream);
int gsb_gbuffer_bwd(const float* rast, const float* v_pos, const float* v_nrm, const float* msdf, const int32_t* tris, int64_t n_batch,
                    int64_t H, int64_t W, int64_t n_verts, const float* g_pos, const float* g_nrm, const float* g_geo_nrm,
                    const float* g_msdf_img, float* g_v_pos, float* g_v_nrm, float* g_msdf, float* g_rast, void* stream);
int gsb_compose_fwd(const void* const* in12, const float* bg, int bg_batched, int64_t n_batch, int64_t H, int64_t W, int mode, int composite,
                    void* const* out12, void* stream);
int gsb_compose_bwd(const void* const* in12, int64_t n_batch, int64_t H, int64_t W, int mode, int composite, const void* const* gin9,
                    void* const* gout7, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Silhouette antialiasing with vertex gradients (csrc/antialias.cu; stands where the reference calls nvdiffrast's
 * dr.antialias, render/render.py:352-359 -- third-party and absent here, own implementation of its documented algorithm,
 * parity unpinned).  Analyse once per frame (edge hash of the mesh + one work item per silhouette crossing between adjacent
 * pixels), then replay the items for every buffer, forward and backward.
 *   hash_ws: gsb_antialias_hash_slots(n_faces) * 16 bytes; items: item_cap * gsb_antialias_item_bytes() bytes, item_cap =
 *   2 * B * H * W always suffices; n_items: device int32.  fwd: `out` holds a copy of color [B,H,W,C] on entry.
 *   bwd: g_color (or NULL) holds a copy of g_out on entry; g_clip float[B,V,4] (or NULL) is accumulated.
 * ---------------------------------------------------------------------------------------------- */
int64_t gsb_antialias_hash_slots(int64_t n_faces);
size_t gsb_antialias_item_bytes(void);
int gsb_antialias_analyse(const float* rast, const float* clip, const int32_t* tris, int64_t n_batch, int64_t H, int64_t W, int64_t n_verts,
                          int64_t n_faces, void* hash_ws, void* items, int32_t* n_items, int64_t item_cap, void* stream);
int gsb_antialias_fwd(const float* color, const void* items, const int32_t* n_items, int64_t item_cap, int64_t n_channels, float* out,
                      void* stream);
int gsb_antialias_bwd(const float* color, const float* g_out, const void* items, const int32_t* n_items, int64_t item_cap, int64_t n_channels,
                      const float* clip, int64_t n_verts, int64_t H, int64_t W, float* g_color, float* g_clip, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pointwise BSDF operators of render/renderutils (csrc/bsdf_ops.cu).  Each pair replaces the plugin functions of the same name
 * that the reference binds at render/renderutils/c_src/torch_bindings.cpp:1034-1061 (`lambert_fwd/bwd`, `frostbite_fwd/bwd`,
 * `pbr_specular_fwd/bwd`, `pbr_bsdf_fwd/bwd`, `fresnel_shlick_fwd/bwd`, `ndf_ggx_fwd/bwd`, `lambda_ggx_fwd/bwd`,
 * `masking_smith_fwd/bwd`, `xfm_fwd/bwd` with isPoints = false; Python side render/renderutils/ops.py:91-390, :540-556).
 * All arrays are dense device arrays of n elements: "3" = float[n,3], "1" = float[n]; the caller broadcasts.  The bwd entry points
 * write every gradient array (no accumulation).
 *   fresnel_shlick: f0 3, f90 3, cos_theta 1 -> 3          ndf_ggx / lambda_ggx: alpha_sqr 1, cos_theta 1 -> 1
 *   masking_smith:  alpha_sqr 1, cos_i 1, cos_o 1 -> 1     lambert: nrm 3, wi 3 -> 1
 *   frostbite:      nrm 3, wi 3, wo 3, linear_roughness 1 -> 1
 *   pbr_specular:   col 3, nrm 3, wo 3, wi 3, alpha 1 -> 3
 *   pbr_bsdf:       HOST array of 6 device pointers (kd, arm, pos, nrm, view_pos, light_pos), all 3 -> 3; bsdf 0 lambert, 1 frostbite
 *   xfm_vectors:    vectors [1,N,3] (vectors_batched = 0) or [B,N,3], matrix [B,4,4] -> out [B,N,3] = matrix[b][:3,:3] * v
 * ---------------------------------------------------------------------------------------------- */
int gsb_fresnel_shlick_fwd(const float* f0, const float* f90, const float* cos_theta, int64_t n, float* out, void* stream);
int gsb_fresnel_shlick_bwd(const float* f0, const float* f90, const float* cos_theta, const float* g_out, int64_t n, float* g_f0, float* g_f90,
                           float* g_cos_theta, void* stream);
int gsb_ndf_ggx_fwd(const float* alpha_sqr, const float* cos_theta, int64_t n, float* out, void* stream);
int gsb_ndf_ggx_bwd(const float* alpha_sqr, const float* cos_theta, const float* g_out, int64_t n, float* g_alpha_sqr, float* g_cos_theta,
                    void* stream);
int gsb_lambda_ggx_fwd(const float* alpha_sqr, const float* cos_theta, int64_t n, float* out, void* stream);
int gsb_lambda_ggx_bwd(const float* alpha_sqr, const float* cos_theta, const float* g_out, int64_t n, float* g_alpha_sqr, float* g_cos_theta,
                       void* stream);
int gsb_masking_smith_fwd(const float* alpha_sqr, const float* cos_i, const float* cos_o, int64_t n, float* out, void* stream);
int gsb_masking_smith_bwd(const float* alpha_sqr, const float* cos_i, const float* cos_o, const float* g_out, int64_t n, float* g_alpha_sqr,
                          float* g_cos_i, float* g_cos_o, void* stream);
int gsb_lambert_fwd(const float* nrm, const float* wi, int64_t n, float* out, void* stream);
int gsb_lambert_bwd(const float* nrm, const float* wi, const float* g_out, int64_t n, float* g_nrm, float* g_wi, void* stream);
int gsb_frostbite_fwd(const float* nrm, const float* wi, const float* wo, const float* linear_roughness, int64_t n, float* out, void* stream);
int gsb_frostbite_bwd(const float* nrm, const float* wi, const float* wo, const float* linear_roughness, const float* g_out, int64_t n,
                      float* g_nrm, float* g_wi, float* g_wo, float* g_linear_roughness, void* stream);
int gsb_pbr_specular_fwd(const float* col, const float* nrm, const float* wo, const float* wi, const float* alpha, float min_roughness, int64_t n,
                         float* out, void* stream);
int gsb_pbr_specular_bwd(const float* col, const float* nrm, const float* wo, const float* wi, const float* alpha, float min_roughness,
                         const float* g_out, int64_t n, float* g_col, float* g_nrm, float* g_wo, float* g_wi, float* g_alpha, void* stream);
int gsb_pbr_bsdf_fwd(const float* const* inputs6_host, float min_roughness, int bsdf, int64_t n, float* out, void* stream);
int gsb_pbr_bsdf_bwd(const float* const* inputs6_host, float min_roughness, int bsdf, const float* g_out, int64_t n, float* const* g_inputs6_host,
                     void* stream);
int gsb_xfm_vectors_fwd(const float* vectors, const float* matrix, int64_t n_batch, int64_t n_vectors, int vectors_batched, float* out,
                        void* stream);
int gsb_xfm_vectors_bwd(const float* matrix, const float* g_out, int64_t n_batch, int64_t n_vectors, int vectors_batched, float* g_vectors,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSHELL_B200_H */
