// TEST INFRASTRUCTURE.  Thread-block execution for the host build of kernels that use shared memory, barriers and warp
// collectives (included by cuda_host/cuda_runtime.h when the translation unit is compiled with -DGSB_HOST_BLOCKS).
//
// One block at a time.  Every thread of the block is a fiber (ucontext) with its own stack; a fiber runs until it finishes or
// reaches a point where it must wait for other threads -- __syncthreads() or a *_sync warp collective -- and then switches back
// to the scheduler, which resumes the fibers whose wait is over, in thread order.  `__shared__` is `static` (one block at a time
// makes that the block's memory), `extern __shared__` arrays come from a per-launch buffer filled with NaN patterns.
// Deterministic, single OS thread; a block in which no fiber can make progress (a barrier not reached by every live thread, a
// collective whose mask names a lane that never arrives) aborts with a message instead of hanging.
// Not modelled: memory ordering (threads interleave only at the wait points above), divergence inside a warp (a collective is a
// rendezvous of the lanes its mask names that have not exited), clusters, TMA, tensor cores.
#pragma once
#include <sys/mman.h>
#include <ucontext.h>

#include <cstdio>

#if defined(__SANITIZE_ADDRESS__)
extern "C" void __sanitizer_start_switch_fiber(void** fake_stack_save, const void* bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void* fake_stack_save, const void** bottom_old, size_t* size_old);
#endif

namespace gsb_host {

constexpr size_t kFiberStack = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Warp {
  unsigned alive = 0;                 // lanes that exist and have not returned
  unsigned arrived[2] = {0, 0};
  unsigned consumed[2] = {0, 0};
  unsigned long long value[2][32];
};

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  void* fake_stack = nullptr;
  bool done = true;
  bool at_barrier = false;
  unsigned barrier_gen = 0;
  bool at_collective = false;
  unsigned collective_need = 0;
  int collective_slot = 0;
  unsigned collective_seq = 0;
  uint3 tid;
};

struct BlockState {
  Fiber fibers[kMaxThreads];
  Warp warps[kMaxThreads / 32];
  ucontext_t scheduler;
  void* scheduler_fake_stack = nullptr;
  const void* scheduler_stack_bottom = nullptr;
  size_t scheduler_stack_size = 0;
  int n_threads = 0, current = -1, alive = 0;
  unsigned barrier_gen = 0;
  int barrier_count = 0;
  void (*call)(void*) = nullptr;
  void* call_ctx = nullptr;
  std::vector<unsigned char> dynamic_shared;
};

inline BlockState& block_state() {
  static BlockState* s = new BlockState();
  return *s;
}

inline void* dynamic_shared() { return block_state().dynamic_shared.data(); }

inline void switch_to_scheduler() {
  BlockState& b = block_state();
  Fiber& f = b.fibers[b.current];
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_start_switch_fiber(f.done ? nullptr : &f.fake_stack, b.scheduler_stack_bottom, b.scheduler_stack_size);
#endif
  swapcontext(&f.ctx, &b.scheduler);
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_finish_switch_fiber(f.fake_stack, &b.scheduler_stack_bottom, &b.scheduler_stack_size);
#endif
}

inline void fiber_entry() {
  BlockState& b = block_state();
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_finish_switch_fiber(nullptr, &b.scheduler_stack_bottom, &b.scheduler_stack_size);
#endif
  b.call(b.call_ctx);
  Fiber& f = b.fibers[b.current];
  f.done = true;
  b.alive--;
  b.warps[b.current / 32].alive &= ~(1u << (b.current % 32));
  switch_to_scheduler();
}

inline bool fiber_ready(BlockState& b, int t) {
  Fiber& f = b.fibers[t];
  if (f.done) return false;
  if (f.at_barrier) return f.barrier_gen != b.barrier_gen;
  if (f.at_collective) {
    const Warp& w = b.warps[t / 32];
    const unsigned need = f.collective_need & w.alive;
    return (w.arrived[f.collective_slot] & need) == need;
  }
  return true;
}

inline void run_block(dim3 block, void (*call)(void*), void* ctx) {
  BlockState& b = block_state();
  const int n = (int)(block.x * block.y * block.z);
  if (n > kMaxThreads) { std::fprintf(stderr, "gsb_host: block of %d threads\n", n); std::abort(); }
  b.n_threads = n;
  b.alive = n;
  b.barrier_count = 0;
  b.call = call;
  b.call_ctx = ctx;
  for (int w = 0; w < (n + 31) / 32; ++w) {
    b.warps[w] = Warp();
    const int lanes = std::min(32, n - 32 * w);
    b.warps[w].alive = lanes == 32 ? 0xFFFFFFFFu : ((1u << lanes) - 1u);
  }
  for (int t = 0; t < n; ++t) {
    Fiber& f = b.fibers[t];
    if (!f.stack) {
      f.stack = (char*)mmap(nullptr, kFiberStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (f.stack == (char*)MAP_FAILED) { std::perror("mmap"); std::abort(); }
    }
    f.done = false;
    f.at_barrier = f.at_collective = false;
    f.collective_seq = 0;
    f.fake_stack = nullptr;
    f.tid.x = (unsigned)(t % block.x); f.tid.y = (unsigned)((t / block.x) % block.y); f.tid.z = (unsigned)(t / (block.x * block.y));
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kFiberStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  while (b.alive > 0) {
    if (b.barrier_count > 0 && b.barrier_count >= b.alive) { b.barrier_gen++; b.barrier_count = 0; }
    bool ran = false;
    for (int k = 0; k < n; ++k) {
      const int t = (gsb_host_thread_order_seed & 1u) ? n - 1 - k : k;      // odd seeds: the threads of a block in descending order
      if (!fiber_ready(b, t)) continue;
      ran = true;
      b.current = t;
      threadIdx = b.fibers[t].tid;
#if defined(__SANITIZE_ADDRESS__)
      __sanitizer_start_switch_fiber(&b.scheduler_fake_stack, b.fibers[t].stack, kFiberStack);
#endif
      swapcontext(&b.scheduler, &b.fibers[t].ctx);
#if defined(__SANITIZE_ADDRESS__)
      __sanitizer_finish_switch_fiber(b.scheduler_fake_stack, nullptr, nullptr);
#endif
      if (b.barrier_count > 0 && b.barrier_count >= b.alive) { b.barrier_gen++; b.barrier_count = 0; }
    }
    if (!ran && b.alive > 0) {
      std::fprintf(stderr, "gsb_host: deadlock in block (%u,%u,%u): %d live threads, %d at the barrier\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   b.alive, b.barrier_count);
      std::abort();
    }
  }
}

inline void sync_threads() {
  BlockState& b = block_state();
  Fiber& f = b.fibers[b.current];
  f.at_barrier = true;
  f.barrier_gen = b.barrier_gen;
  b.barrier_count++;
  switch_to_scheduler();
  f.at_barrier = false;
}

// Rendezvous of the lanes `mask` names (that have not exited): every lane contributes `v`, out[l] = lane l's value; returns the
// set of lanes that took part.
inline unsigned warp_exchange(unsigned mask, unsigned long long v, unsigned long long* out) {
  BlockState& b = block_state();
  const int t = b.current, lane = t % 32;
  Fiber& f = b.fibers[t];
  Warp& w = b.warps[t / 32];
  const int slot = (int)(f.collective_seq++ & 1u);
  w.value[slot][lane] = v;
  w.arrived[slot] |= 1u << lane;
  f.collective_need = mask;
  f.collective_slot = slot;
  if ((w.arrived[slot] & (mask & w.alive)) != (mask & w.alive)) {
    f.at_collective = true;
    switch_to_scheduler();
    f.at_collective = false;
  }
  const unsigned part = w.arrived[slot] & mask;
  for (int l = 0; l < 32; ++l) out[l] = w.value[slot][l];
  w.consumed[slot] |= 1u << lane;
  if ((w.consumed[slot] & part) == part) { w.arrived[slot] &= ~part; w.consumed[slot] &= ~part; }
  return part;
}

inline int lane_id() { return block_state().current % 32; }

template <class T> inline unsigned long long to_bits(T v) { unsigned long long u = 0; std::memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(unsigned long long u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }

template <class T> inline T shfl_from(unsigned mask, T v, int src_lane, bool in_range) {
  unsigned long long all[32];
  const unsigned part = warp_exchange(mask, to_bits(v), all);
  if (!in_range || !((part >> src_lane) & 1u)) return v;
  return from_bits<T>(all[src_lane]);
}

template <class F> inline void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
  BlockState& b = block_state();
  // launch limits of the device (sm_100): a configuration the hardware would refuse must not pass here
  if ((long long)block.x * block.y * block.z > 1024 || block.x > 1024 || block.y > 1024 || block.z > 64 || grid.y > 65535u || grid.z > 65535u ||
      grid.x > 2147483647u || smem > 227u * 1024u || grid.x == 0 || grid.y == 0 || grid.z == 0 || block.x * block.y * block.z == 0) {
    std::fprintf(stderr, "gsb_host: launch configuration the device refuses: grid (%u,%u,%u) block (%u,%u,%u) dynamic shared %zu B\n", grid.x, grid.y,
                 grid.z, block.x, block.y, block.z, smem);
    std::abort();
  }
  blockDim = block;
  gridDim = grid;
  b.dynamic_shared.assign(smem + 16, 0xFF);
  const long long n_blocks = (long long)grid.x * grid.y * grid.z;
  std::vector<long long> order((size_t)n_blocks);
  std::iota(order.begin(), order.end(), 0LL);
  if (gsb_host_thread_order_seed) {
    unsigned long long s = gsb_host_thread_order_seed * 0x9E3779B97F4A7C15ull + 1;
    for (long long i = n_blocks - 1; i > 0; --i) {
      s += 0x9E3779B97F4A7C15ull;
      unsigned long long z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
      std::swap(order[(size_t)i], order[(size_t)(z % (unsigned long long)(i + 1))]);
    }
  }
  using Body = typename std::remove_reference<F>::type;
  for (long long blk : order) {
    blockIdx.x = (unsigned)(blk % grid.x); blockIdx.y = (unsigned)((blk / grid.x) % grid.y); blockIdx.z = (unsigned)(blk / ((long long)grid.x * grid.y));
    std::fill(b.dynamic_shared.begin(), b.dynamic_shared.end(), (unsigned char)0xFF);
    run_block(block, [](void* p) { (*(Body*)p)(); }, (void*)&body);
  }
}
template <class G, class B, class F> inline void launch(G grid, B block, size_t smem, F&& body) { launch(as_dim3(grid), as_dim3(block), smem, body); }

}  // namespace gsb_host

#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
inline void __syncthreads() { gsb_host::sync_threads(); }
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { unsigned long long all[32]; gsb_host::warp_exchange(mask, 0, all); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __trap() { std::fprintf(stderr, "gsb_host: __trap()\n"); std::abort(); }
inline unsigned __activemask() { unsigned long long all[32]; return gsb_host::warp_exchange(0xFFFFFFFFu, 0, all); }
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = gsb_host::lane_id(), base = lane & ~(width - 1);
  return gsb_host::shfl_from(mask, v, base + (src & (width - 1)), true);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = gsb_host::lane_id(), base = lane & ~(width - 1), src = lane - (int)delta;
  return gsb_host::shfl_from(mask, v, src < base ? lane : src, src >= base);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = gsb_host::lane_id(), base = lane & ~(width - 1), src = lane + (int)delta;
  return gsb_host::shfl_from(mask, v, src >= base + width ? lane : src, src < base + width);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
  const int lane = gsb_host::lane_id(), base = lane & ~(width - 1), src = lane ^ lane_mask;
  return gsb_host::shfl_from(mask, v, src >= base + width ? lane : src, src < base + width);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned long long all[32];
  const unsigned part = gsb_host::warp_exchange(mask, pred ? 1ull : 0ull, all);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) if (((part >> l) & 1u) && all[l]) r |= 1u << l;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
  unsigned long long all[32];
  const unsigned part = gsb_host::warp_exchange(mask, pred ? 1ull : 0ull, all);
  for (int l = 0; l < 32; ++l) if (((part >> l) & 1u) && !all[l]) return 0;
  return 1;
}
template <class T> inline T gsb_host_reduce_add(unsigned mask, T v) {
  unsigned long long all[32];
  const unsigned part = gsb_host::warp_exchange(mask, gsb_host::to_bits(v), all);
  T r = 0;
  for (int l = 0; l < 32; ++l) if ((part >> l) & 1u) r += gsb_host::from_bits<T>(all[l]);
  return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return gsb_host_reduce_add(mask, v); }
inline int __reduce_add_sync(unsigned mask, int v) { return gsb_host_reduce_add(mask, v); }
