// TEST INFRASTRUCTURE: kernels with known answers for the thread-block emulator itself (tests/native/cuda_host/block_emulator.h).
// Plain CUDA -- this file also compiles with nvcc; tests/test_block_emulator_cpu.py runs it through host_kernels.build(blocks=True).
#include <cuda_runtime.h>
#include <stdint.h>
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

namespace {
constexpr unsigned kFull = 0xFFFFFFFFu;

// inclusive scan of one 256-element tile per block: shuffles inside a warp, shared memory across warps, two barriers
__global__ void k_block_scan(const int* in, int* out) {
  __shared__ int warp_sum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, i = blockIdx.x * 256 + threadIdx.x;
  int v = in[i];
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += y;
  }
  if (lane == 31) warp_sum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = lane < 8 ? warp_sum[lane] : 0;
    for (int o = 1; o < 8; o <<= 1) {
      const int y = __shfl_up_sync(kFull, w, o);
      if (lane >= o) w += y;
    }
    if (lane < 8) warp_sum[lane] = w;
  }
  __syncthreads();
  out[i] = v + (warp ? warp_sum[warp - 1] : 0);
}

// order-preserving compaction of the positive entries of a tile: ballot + popc ranks, a shared counter per warp
__global__ void k_compact(const int* in, int* out, int* count) {
  __shared__ int base[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = in[blockIdx.x * 256 + threadIdx.x];
  const unsigned b = __ballot_sync(kFull, v > 0);
  if (lane == 0) base[warp] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < 8; ++w) { const int c = base[w]; base[w] = run; run += c; }
    count[blockIdx.x] = run;
  }
  __syncthreads();
  if (v > 0) out[blockIdx.x * 256 + base[warp] + __popc(b & ((1u << lane) - 1u))] = v;
}

// reversal of a tile through dynamic shared memory
__global__ void k_reverse(const float* in, float* out, int n) {
  extern __shared__ float tile[];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  tile[threadIdx.x] = i < n ? in[i] : 0.f;
  __syncthreads();
  const int j = blockDim.x - 1 - threadIdx.x;
  if (blockIdx.x * blockDim.x + j < n || true) out[i] = tile[j];
}

// butterfly sum, any / all, reduce_add, a 16-wide segmented broadcast
__global__ void k_warp_ops(const int* in, int* sum, int* any_neg, int* all_pos, int* redux, int* seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int v = in[i];
  int s = v;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
  sum[i] = s;
  any_neg[i] = __any_sync(kFull, v < 0);
  all_pos[i] = __all_sync(kFull, v > 0);
  redux[i] = __reduce_add_sync(kFull, v);
  seg[i] = __shfl_sync(kFull, v, 3, 16);            // lane 3 of the thread's 16-lane segment
}

// 3-D indices
__global__ void k_indices(int* out) {
  const int b = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int t = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
  const int per = blockDim.x * blockDim.y * blockDim.z;
  out[(b * per + t) * 2] = b;
  out[(b * per + t) * 2 + 1] = t;
}

// threads with a negative input leave first; the rest form a coalesced group and scan their values
__global__ void k_coalesced(const int* in, int* rank, int* excl, int* size) {
  namespace cg = cooperative_groups;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  rank[i] = excl[i] = size[i] = -1;
  if (in[i] < 0) return;
  cg::coalesced_group g = cg::coalesced_threads();
  rank[i] = (int)g.thread_rank();
  size[i] = (int)g.size();
  excl[i] = cg::exclusive_scan(g, in[i]);
}

// ping-pong through shared memory: every round depends on the neighbour's value of the round before
__global__ void k_rounds(int* out, int rounds) {
  __shared__ int a[128], b[128];
  a[threadIdx.x] = threadIdx.x;
  __syncthreads();
  for (int r = 0; r < rounds; ++r) {
    b[threadIdx.x] = a[(threadIdx.x + 1) & 127] + 1;
    __syncthreads();
    a[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
  }
  out[blockIdx.x * 128 + threadIdx.x] = a[threadIdx.x];
}
}  // namespace

extern "C" {
int st_block_scan(const int* in, int* out, int n_blocks) { k_block_scan<<<n_blocks, 256>>>(in, out); return (int)cudaGetLastError(); }
int st_compact(const int* in, int* out, int* count, int n_blocks) { k_compact<<<n_blocks, 256>>>(in, out, count); return (int)cudaGetLastError(); }
int st_reverse(const float* in, float* out, int n, int block) {
  k_reverse<<<(n + block - 1) / block, block, block * sizeof(float)>>>(in, out, n);
  return (int)cudaGetLastError();
}
int st_warp_ops(const int* in, int* sum, int* any_neg, int* all_pos, int* redux, int* seg, int n_blocks) {
  k_warp_ops<<<n_blocks, 64>>>(in, sum, any_neg, all_pos, redux, seg);
  return (int)cudaGetLastError();
}
int st_indices(int* out) { k_indices<<<dim3(3, 2, 2), dim3(4, 3, 2)>>>(out); return (int)cudaGetLastError(); }
int st_coalesced(const int* in, int* rank, int* excl, int* size, int n_blocks) {
  k_coalesced<<<n_blocks, 64>>>(in, rank, excl, size);
  return (int)cudaGetLastError();
}
int st_rounds(int* out, int rounds, int n_blocks) { k_rounds<<<n_blocks, 128>>>(out, rounds); return (int)cudaGetLastError(); }
}
