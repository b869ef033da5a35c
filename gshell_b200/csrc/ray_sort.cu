// Coherence sort of the shadow-ray list (keys built by k_env_shade<GEN>): rays are ordered by direction bucket, then by
// the Morton code of where they pierce the grid's mid-plane, so that the rays a persistent warp (and the whole GPU) works on
// at any moment form a narrow beam through the occluder grid -- their cells and triangle records stay in L1/L2 instead of
// being re-fetched from HBM (the unsorted kernel read 705 GB per launch for ~10 GB of algorithmic bytes, profiles/r1f).
// The sort itself is CUB's device radix sort (library code, like a plain cuBLAS call): 32-bit keys + 32-bit ray indices.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gshell_b200.h"

extern "C" {

size_t gsb_ray_sort_temp_bytes(int64_t max_items) {
  cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int)max_items, 0, 32, (cudaStream_t)0);
  return bytes + 256;
}

/* keys/idx: ping-pong buffers [2][n] (first half holds the input); returns in *sorted_idx_offset 0 or n = which half of
 * idx holds the sorted permutation. */
int gsb_ray_sort(uint32_t* keys, uint32_t* idx, int64_t capacity, int64_t n, void* temp, size_t temp_bytes,
                 int64_t* sorted_idx_offset, void* stream) {
  if (n <= 0) { *sorted_idx_offset = 0; return 0; }
  cub::DoubleBuffer<uint32_t> k(keys, keys + capacity), v(idx, idx + capacity);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, (int)n, 0, 32, (cudaStream_t)stream);
  *sorted_idx_offset = (v.Current() == idx) ? 0 : capacity;
  return (int)e;
}

}  // extern "C"
