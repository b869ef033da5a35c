// TEST INFRASTRUCTURE.  A host stand-in for the few pieces of <cuda_runtime.h> that the product's "one independent thread per
// element" kernels use, so that the UNMODIFIED kernel source of such a translation unit (csrc/mesh_ops.cu, csrc/tangents.cu)
// compiles with g++ and runs on the CPU inside the `-m "not gpu"` suite (tests/native/host_kernels.py rewrites the
// `kernel<<<grid, block, smem, stream>>>(args)` launches to gsb_host::launch before compiling).
// Default mode: no shared memory, no barriers, no warp intrinsics -- a kernel that needs them does not compile, which is the
// intended failure.  Threads run one after the other, in an order drawn from gsb_host_thread_order_seed (0 = ascending), so that
// sums accumulated with atomicAdd see different summation orders, as on the GPU.
// With -DGSB_HOST_BLOCKS (host_kernels.build(..., blocks=True)) block_emulator.h adds thread blocks: one fiber per thread,
// `__shared__`, __syncthreads() and the *_sync warp collectives as rendezvous points; blocks run one after the other.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

// ---- vector types ------------------------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int3 make_int3(int x, int y, int z) { return int3{x, y, z}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline uint3 threadIdx, blockIdx;        // one instance for all translation units of the library
inline dim3 blockDim, gridDim;
#define __grid_constant__
#define __constant__ static

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
typedef void* cudaStream_t;

inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaErrorMemoryAllocation = 2 };
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMemcpyFromSymbol(void* d, const T& symbol, size_t n) { std::memcpy(d, &symbol, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMemcpyToSymbol(T& symbol, const void* s, size_t n) { std::memcpy(&symbol, s, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)std::malloc(n); return *p ? cudaSuccess : 2; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)std::malloc(n); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
typedef int cudaEvent_t;
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = 0; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }

template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
// atomics: threads run one after the other, so read-modify-write is atomic by construction
template <class T> inline T gsb_host_rmw_add(T* p, T v) { T o = *p; *p = o + v; return o; }
inline float atomicAdd(float* p, float v) { return gsb_host_rmw_add(p, v); }
inline double atomicAdd(double* p, double v) { return gsb_host_rmw_add(p, v); }
inline int atomicAdd(int* p, int v) { return gsb_host_rmw_add(p, v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return gsb_host_rmw_add(p, v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return gsb_host_rmw_add(p, v); }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T expect, T v) { T o = *p; if (o == expect) *p = v; return o; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
// separately rounded operations: the translation unit is compiled with -ffp-contract=off
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float a) { return std::exp(a); }
inline float __logf(float a) { return std::log(a); }
inline float __powf(float a, float b) { return std::pow(a, b); }
inline float __sinf(float a) { return std::sin(a); }
inline float __cosf(float a) { return std::cos(a); }
inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
using std::max;
using std::min;

extern "C" unsigned gsb_host_thread_order_seed;

namespace gsb_host {
// vector reductions the kernels issue as inline PTX (`red.global.add.v2/v4.f32`); host_kernels.py rewrites the asm statement to this
inline void red_add(float* p, float a, float b) { p[0] += a; p[1] += b; }
inline void red_add(float* p, float a, float b, float c, float d) { p[0] += a; p[1] += b; p[2] += c; p[3] += d; }

inline dim3 as_dim3(dim3 d) { return d; }
inline dim3 as_dim3(long long n) { return dim3((unsigned)n); }

#ifndef GSB_HOST_BLOCKS
template <class F> inline void launch(dim3 grid, dim3 block, size_t /*dynamic shared memory: none without GSB_HOST_BLOCKS*/, F&& body) {
  const long long per_block = (long long)block.x * block.y * block.z;
  const long long n = (long long)grid.x * grid.y * grid.z * per_block;
  blockDim = block;
  gridDim = grid;
  std::vector<long long> order((size_t)n);
  std::iota(order.begin(), order.end(), 0LL);
  if (gsb_host_thread_order_seed) {
    unsigned long long s = gsb_host_thread_order_seed * 0x9E3779B97F4A7C15ull + 1;
    for (long long i = n - 1; i > 0; --i) {                  // Fisher-Yates with a splitmix-style stream
      s += 0x9E3779B97F4A7C15ull;
      unsigned long long z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
      std::swap(order[(size_t)i], order[(size_t)(z % (unsigned long long)(i + 1))]);
    }
  }
  for (long long t : order) {
    long long b = t / per_block, l = t % per_block;
    blockIdx.x = (unsigned)(b % grid.x); blockIdx.y = (unsigned)((b / grid.x) % grid.y); blockIdx.z = (unsigned)(b / ((long long)grid.x * grid.y));
    threadIdx.x = (unsigned)(l % block.x); threadIdx.y = (unsigned)((l / block.x) % block.y); threadIdx.z = (unsigned)(l / ((long long)block.x * block.y));
    body();
  }
}
template <class G, class B, class F> inline void launch(G grid, B block, size_t smem, F&& body) { launch(as_dim3(grid), as_dim3(block), smem, body); }
#endif
}  // namespace gsb_host

#ifdef GSB_HOST_BLOCKS
#include "block_emulator.h"
#endif
