// TEST INFRASTRUCTURE (oracle/): host driver around the reference's UNMODIFIED envsampling/kernel.cu, compiled for the CPU by
// oracle/build_ref.py with the shims in oracle/ref_shim/.  It plays the role of torch_bindings.cpp:123-272 (fill the
// EnvSamplingParams accessors, zero-initialised outputs, launch over (W, H, B)) and of the OptiX runtime (launch index,
// any-hit query).  The reference source is included from where it lies under /root/reference; nothing is copied.
#include REF_KERNEL_CU

#include <vector>

thread_local uint3 g_ref_launch_index;
uint3 g_ref_launch_dims;

namespace {
const float* g_verts = nullptr;
const int* g_tris = nullptr;
int g_n_tris = 0;

// layout of PackedTensorAccessor32<T, N> (accessor.h: data pointer, int32 sizes[N], int32 strides[N])
template <class T, int N>
struct Packed { T* data; int32_t sizes[N]; int32_t strides[N]; };

template <class T, int N, class Dst>
void bind(Dst& dst, const T* data, const int (&sizes)[N]) {
  static_assert(sizeof(Dst) == sizeof(Packed<T, N>), "PackedTensorAccessor32 layout changed");
  Packed<T, N> p;
  p.data = const_cast<T*>(data);
  int stride = 1;
  for (int i = N - 1; i >= 0; --i) { p.sizes[i] = sizes[i]; p.strides[i] = stride; stride *= sizes[i]; }
  memcpy((void*)&dst, &p, sizeof(p));
}
}  // namespace

// two-sided Moeller-Trumbore in double precision over every triangle: the geometric ground truth of "any hit in (tmin, tmax)"
bool ref_any_hit(float3 o, float3 d, float tmin, float tmax) {
  for (int f = 0; f < g_n_tris; ++f) {
    const float* a = g_verts + 3 * g_tris[3 * f];
    const float* b = g_verts + 3 * g_tris[3 * f + 1];
    const float* c = g_verts + 3 * g_tris[3 * f + 2];
    const double e1[3] = {(double)b[0] - a[0], (double)b[1] - a[1], (double)b[2] - a[2]};
    const double e2[3] = {(double)c[0] - a[0], (double)c[1] - a[1], (double)c[2] - a[2]};
    const double p[3] = {d.y * e2[2] - d.z * e2[1], d.z * e2[0] - d.x * e2[2], d.x * e2[1] - d.y * e2[0]};
    const double det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (det == 0.0) continue;
    const double inv = 1.0 / det;
    const double t[3] = {(double)o.x - a[0], (double)o.y - a[1], (double)o.z - a[2]};
    const double u = (t[0] * p[0] + t[1] * p[1] + t[2] * p[2]) * inv;
    if (u < 0.0 || u > 1.0) continue;
    const double q[3] = {t[1] * e1[2] - t[2] * e1[1], t[2] * e1[0] - t[0] * e1[2], t[0] * e1[1] - t[1] * e1[0]};
    const double v = (d.x * q[0] + d.y * q[1] + d.z * q[2]) * inv;
    if (v < 0.0 || u + v > 1.0) continue;
    const double tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
    if (tt > tmin && tt < tmax) return true;
  }
  return false;
}

extern "C" {

// All tensors dense fp32 NHWC; view_pos is [B,1,1,3] (broadcast inside the reference through its size-1 checks); rows is the
// dense [lh] vector.  backward == 0: writes diff, spec ([B,H,W,3], zero-initialised here).  backward == 1: reads diff_grad /
// spec_grad and writes the five zero-initialised gradient tensors.  verts/tris (may be null / 0) define the occluders.
void ref_env_shade(int backward, int B, int H, int W, const float* mask, const float* ro, const float* pos, const float* nrm,
                   const float* view_pos, const float* kd, const float* ks, const float* light, const float* pdf,
                   const float* rows, const float* cols, const int* perms, int n_perms, int lh, int lw, unsigned bsdf,
                   unsigned n_samples_x, unsigned rnd_seed, float shadow_scale, const float* verts, const int* tris, int n_tris,
                   float* diff, float* spec, const float* diff_grad, const float* spec_grad, float* pos_grad, float* nrm_grad,
                   float* kd_grad, float* ks_grad, float* light_grad) {
  g_verts = verts; g_tris = tris; g_n_tris = n_tris;
  const int full[4] = {B, H, W, 3}, m3[3] = {B, H, W}, vp[4] = {B, 1, 1, 3}, l3[3] = {lh, lw, 3}, l2[2] = {lh, lw}, l1[1] = {lh};
  const int pm[2] = {n_perms, (int)(n_samples_x * n_samples_x)};
  const size_t npx = (size_t)B * H * W * 3;
  bind<float, 3>(params.mask, mask, m3);
  bind<float, 4>(params.ro, ro, full);
  bind<float, 4>(params.gb_pos, pos, full);
  bind<float, 4>(params.gb_normal, nrm, full);
  bind<float, 4>(params.gb_view_pos, view_pos, vp);
  bind<float, 4>(params.gb_kd, kd, full);
  bind<float, 4>(params.gb_ks, ks, full);
  bind<float, 3>(params.light, light, l3);
  bind<float, 2>(params.pdf, pdf, l2);
  bind<float, 1>(params.rows, rows, l1);
  bind<float, 2>(params.cols, cols, l2);
  bind<int, 2>(params.perms, perms, pm);
  params.handle = 0;
  params.BSDF = bsdf;
  params.n_samples_x = n_samples_x;
  params.rnd_seed = rnd_seed;
  params.backward = backward ? 1u : 0u;
  params.shadow_scale = shadow_scale;
  if (!backward) {
    memset(diff, 0, npx * sizeof(float)); memset(spec, 0, npx * sizeof(float));
    bind<float, 4>(params.diff, diff, full);
    bind<float, 4>(params.spec, spec, full);
  } else {
    bind<float, 4>(params.diff_grad, diff_grad, full);
    bind<float, 4>(params.spec_grad, spec_grad, full);
    float* outs[4] = {pos_grad, nrm_grad, kd_grad, ks_grad};
    for (float* o : outs) memset(o, 0, npx * sizeof(float));
    memset(light_grad, 0, (size_t)lh * lw * 3 * sizeof(float));
    bind<float, 4>(params.gb_pos_grad, pos_grad, full);
    bind<float, 4>(params.gb_normal_grad, nrm_grad, full);
    bind<float, 4>(params.gb_kd_grad, kd_grad, full);
    bind<float, 4>(params.gb_ks_grad, ks_grad, full);
    bind<float, 3>(params.light_grad, light_grad, l3);
  }
  g_ref_launch_dims = make_uint3((unsigned)W, (unsigned)H, (unsigned)B);     // optixLaunch(..., ro.size(2), ro.size(1), ro.size(0))
  // one "launch": every pixel runs the reference's raygen program; rows are spread over the host threads (OpenMP)
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        g_ref_launch_index = make_uint3((unsigned)x, (unsigned)y, (unsigned)b);
        __raygen__rg();
      }
}

}  // extern "C"
