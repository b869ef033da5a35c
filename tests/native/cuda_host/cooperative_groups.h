// TEST INFRASTRUCTURE: the slice of <cooperative_groups.h> the product uses (coalesced_threads + shfl + exclusive_scan), on the
// warp rendezvous of block_emulator.h.  A coalesced group = the lanes of the warp that have not exited when it is formed.
#pragma once
#include <cuda_runtime.h>

namespace cooperative_groups {
class coalesced_group {
 public:
  explicit coalesced_group(unsigned mask) : mask_(mask) {}
  unsigned size() const { return (unsigned)__builtin_popcount(mask_); }
  unsigned thread_rank() const { return (unsigned)__builtin_popcount(mask_ & ((1u << gsb_host::lane_id()) - 1u)); }
  unsigned mask() const { return mask_; }
  // value of the group's member of rank `rank`
  template <class T> T shfl(T v, unsigned rank) const {
    unsigned long long all[32];
    gsb_host::warp_exchange(mask_, gsb_host::to_bits(v), all);
    unsigned seen = 0;
    for (int l = 0; l < 32; ++l)
      if ((mask_ >> l) & 1u) { if (seen == rank) return gsb_host::from_bits<T>(all[l]); ++seen; }
    return v;
  }
  void sync() const { unsigned long long all[32]; gsb_host::warp_exchange(mask_, 0, all); }

 private:
  unsigned mask_;
};
inline coalesced_group coalesced_threads() { return coalesced_group(__activemask()); }

template <class T> inline T exclusive_scan(const coalesced_group& g, T v) {
  unsigned long long all[32];
  gsb_host::warp_exchange(g.mask(), gsb_host::to_bits(v), all);
  T acc = 0;
  const int lane = gsb_host::lane_id();
  for (int l = 0; l < lane; ++l) if ((g.mask() >> l) & 1u) acc += gsb_host::from_bits<T>(all[l]);
  return acc;
}
}  // namespace cooperative_groups
