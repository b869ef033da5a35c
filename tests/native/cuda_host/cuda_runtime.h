// TEST INFRASTRUCTURE.  A host stand-in for the few pieces of <cuda_runtime.h> that the product's "one independent thread per
// element" kernels use, so that the UNMODIFIED kernel source of such a translation unit (csrc/mesh_ops.cu, csrc/tangents.cu)
// compiles with g++ and runs on the CPU inside the `-m "not gpu"` suite (tests/native/host_kernels.py rewrites the
// `kernel<<<grid, block, smem, stream>>>(args)` launches to gsb_host::launch before compiling).
// Not an emulator: no shared memory, no barriers, no warp intrinsics -- a kernel that needs them does not compile here,
// which is the intended failure.  Threads run one after the other, in an order drawn from gsb_host_thread_order_seed
// (0 = ascending), so that sums accumulated with atomicAdd see different summation orders, as on the GPU.
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

struct gsb_host_uint3 { unsigned x, y, z; };
static gsb_host_uint3 threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
typedef void* cudaStream_t;

inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
enum { cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

template <class T> inline T __ldg(const T* p) { return *p; }
inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
// separately rounded operations: the translation unit is compiled with -ffp-contract=off
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
using std::max;
using std::min;

extern "C" unsigned gsb_host_thread_order_seed;

namespace gsb_host {
template <class F> inline void launch(long long grid, long long block, F&& body) {
  const long long n = grid * block;
  blockDim.x = (unsigned)block; blockDim.y = blockDim.z = 1;
  gridDim.x = (unsigned)grid; gridDim.y = gridDim.z = 1;
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
    blockIdx.x = (unsigned)(t / block); blockIdx.y = blockIdx.z = 0;
    threadIdx.x = (unsigned)(t % block); threadIdx.y = threadIdx.z = 0;
    body();
  }
}
}  // namespace gsb_host
