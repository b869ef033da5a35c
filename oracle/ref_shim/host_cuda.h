/* TEST INFRASTRUCTURE (oracle/): lets g++ compile the reference's UNMODIFIED OptiX device program
 * render/optixutils/c_src/envsampling/kernel.cu (+ common.h, math_utils.h, bsdf.h, accessor.h, params.h) as plain host C++.
 * Force-included before the translation unit (-include).  Nothing here restates reference code: it only supplies what nvcc /
 * NVRTC provide implicitly -- the execution-space qualifiers, the built-in vector types, the overloaded min/max/sincos of the
 * CUDA math library and atomicAdd. */
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* the reference defines M_PI itself as a FLOAT when the compiler has not (NVRTC has not): keep device semantics */
#ifdef M_PI
#undef M_PI
#endif

#define __CUDACC__ 1
#define __device__
#define __host__
#define __global__
#define __constant__
#define __inline__ inline
#define __forceinline__ inline
#define __restrict__

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint3 { unsigned int x, y, z; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint3 make_uint3(unsigned int x, unsigned int y, unsigned int z) { return uint3{x, y, z}; }

/* CUDA's overload set of min / max (same promotion rules as the device math library) */
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline double min(float a, double b) { return fmin((double)a, b); }
static inline double max(float a, double b) { return fmax((double)a, b); }
static inline double min(double a, float b) { return fmin(a, (double)b); }
static inline double max(double a, float b) { return fmax(a, (double)b); }

static inline void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
/* the driver runs pixels on OpenMP threads: the light-gradient accumulation is the only shared write */
static inline float atomicAdd(float* p, float v) {
#pragma omp atomic
  *p += v;
  return 0.f;   /* return value unused by the reference */
}
