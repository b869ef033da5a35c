// Small float3 helpers shared by the G-buffer kernels (own formulation; semantics follow the
// reference's render/optixutils/c_src/math_utils.h:134-198 where a citation says so).
#pragma once
#include <cuda_runtime.h>

namespace gsb {

struct V3 {
  float x, y, z;
};

__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 v3(float s) { return V3{s, s, s}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__device__ __forceinline__ V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float hsum(V3 a) { return a.x + a.y + a.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// adjoint of cross(a, b)
__device__ __forceinline__ void cross_bwd(V3 a, V3 b, V3 g, V3& ga, V3& gb) {
  ga += cross(b, g);
  gb += cross(g, a);
}
// v / |v|, zero for the zero vector (math_utils.h:160-164)
__device__ __forceinline__ V3 normalize0(V3 v) {
  float l = sqrtf(dot(v, v));
  return l > 0.f ? v / l : v3(0.f);
}
// adjoint of normalize0: (g |v|^2 - v (v.g)) / |v|^3
__device__ __forceinline__ V3 normalize0_bwd(V3 v, V3 g) {
  float l2 = dot(v, v);
  if (!(l2 > 0.f)) return v3(0.f);
  float inv3 = 1.f / (l2 * sqrtf(l2));
  float vg = dot(v, g);
  return (g * l2 - v * vg) * inv3;
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
__device__ __forceinline__ float luminance(V3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }

// Orthonormal basis around unit n (Duff et al. 2017; math_utils.h:190-198)
__device__ __forceinline__ void onb(V3 n, V3& b1, V3& b2) {
  float s = copysignf(1.f, n.z);
  float a = -1.f / (s + n.z);
  float b = n.x * n.y * a;
  b1 = V3{1.f + s * n.x * n.x * a, s * b, -s * n.x};
  b2 = V3{b, s + n.y * n.y * a, -n.y};
}

__device__ __forceinline__ V3 ld3(const float* __restrict__ p) { return V3{__ldg(p), __ldg(p + 1), __ldg(p + 2)}; }
__device__ __forceinline__ void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

}  // namespace gsb
