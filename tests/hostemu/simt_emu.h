// simt_emu.h -- TEST INFRASTRUCTURE: runs a CUDA thread block on the host, one coroutine (ucontext) per CUDA thread,
// so that the __global__ functions of svo_kernels.cu -- including their warp-level protocol: __ballot_sync,
// __shfl_sync, __syncwarp, __syncthreads_or, shared-memory queues, atomics -- execute in the CPU test suite.
//
// Semantics: threads of a block run cooperatively on one OS thread and switch only inside a collective, which
// completes when every thread of its group (warp or block) that has not exited has arrived -- the behaviour of
// sm_70+ for the full-mask collectives the kernels use.  `__shared__` becomes `static thread_local` (one block per OS
// thread at a time).  What this cannot model: independent thread scheduling races, memory-model effects, timing.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

namespace simt {

struct Idx3 { unsigned x = 0, y = 0, z = 0; };
struct Dim3 {
  unsigned x = 1, y = 1, z = 1;
  Dim3() {}
  Dim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct Bar { int arrived = 0; unsigned gen = 0; };

struct Block {
  static constexpr int kMaxThreads = 1024;
  static constexpr size_t kStack = 256 * 1024;
  int n = 0, alive = 0, cur = 0;
  ucontext_t main;
  std::vector<ucontext_t> ctx;
  std::vector<char *> stacks;
  std::vector<char> done;
  int warp_alive[kMaxThreads / 32];
  Bar warp_bar[kMaxThreads / 32], cta_bar;
  uint32_t vote[kMaxThreads];
  uint64_t val[kMaxThreads];
  int cta_flag = 0;
  std::function<void()> body;
};

extern thread_local Block *g_block;
extern thread_local Idx3 g_threadIdx, g_blockIdx;
extern thread_local Dim3 g_blockDim, g_gridDim;

inline void yield() {
  Block *b = g_block;
  swapcontext(&b->ctx[(size_t)b->cur], &b->main);
}
// wait until every live thread of the group has arrived; `live` = number of threads of the group that have not exited
template <class Live>
inline void bar_wait(Bar &bar, Live live) {
  bar.arrived++;
  const unsigned g = bar.gen;
  while (bar.gen == g) {
    if (bar.arrived >= live()) { bar.arrived = 0; bar.gen++; break; }
    yield();
  }
}
inline void warp_barrier() {
  Block *b = g_block;
  const int w = b->cur >> 5;
  bar_wait(b->warp_bar[w], [b, w] { return b->warp_alive[w]; });
}
inline void cta_barrier() {
  Block *b = g_block;
  bar_wait(b->cta_bar, [b] { return b->alive; });
}
inline bool lane_alive(int w, int l) {
  Block *b = g_block;
  const int t = w * 32 + l;
  return t < b->n && !b->done[(size_t)t];
}

void run_grid(Dim3 grid, Dim3 block, const std::function<void()> &kernel, int os_threads);

}  // namespace simt

// ---- the CUDA names the kernels use ---------------------------------------------------------------------------
#define threadIdx (simt::g_threadIdx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)
#undef __shared__
#define __shared__ static thread_local
#undef __launch_bounds__
#define __launch_bounds__(...)

static inline uint32_t __ballot_sync(uint32_t, int pred) {
  simt::Block *b = simt::g_block;
  const int w = b->cur >> 5;
  b->vote[b->cur] = pred ? 1u : 0u;
  simt::warp_barrier();
  uint32_t r = 0;
  for (int l = 0; l < 32; l++)
    if (simt::lane_alive(w, l) && b->vote[w * 32 + l]) r |= 1u << l;
  simt::warp_barrier();
  return r;
}
template <class T>
static inline T __shfl_sync(uint32_t, T v, int src) {
  static_assert(sizeof(T) <= 8, "shfl of up to 8 bytes");
  simt::Block *b = simt::g_block;
  const int w = b->cur >> 5;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b->val[b->cur] = raw;
  simt::warp_barrier();
  raw = b->val[w * 32 + (src & 31)];
  simt::warp_barrier();
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_xor_sync(uint32_t, T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "shfl of up to 8 bytes");
  simt::Block *b = simt::g_block;
  const int w = b->cur >> 5, l = b->cur & 31;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b->val[b->cur] = raw;
  simt::warp_barrier();
  raw = b->val[w * 32 + ((l ^ lane_mask) & 31)];
  simt::warp_barrier();
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
static inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) {
  simt::Block *b = simt::g_block;
  const int w = b->cur >> 5;
  b->val[b->cur] = v;
  simt::warp_barrier();
  uint32_t r = 0;
  for (int l = 0; l < 32; l++)
    if (simt::lane_alive(w, l)) r += (uint32_t)b->val[w * 32 + l];
  simt::warp_barrier();
  return r;
}
static inline uint32_t __reduce_min_sync(uint32_t, uint32_t v) {
  simt::Block *b = simt::g_block;
  const int w = b->cur >> 5;
  b->val[b->cur] = v;
  simt::warp_barrier();
  uint32_t r = 0xFFFFFFFFu;
  for (int l = 0; l < 32; l++)
    if (simt::lane_alive(w, l) && (uint32_t)b->val[w * 32 + l] < r) r = (uint32_t)b->val[w * 32 + l];
  simt::warp_barrier();
  return r;
}
static inline uint32_t __reduce_max_sync(uint32_t, uint32_t v) {
  simt::Block *b = simt::g_block;
  const int w = b->cur >> 5;
  b->val[b->cur] = v;
  simt::warp_barrier();
  uint32_t r = 0;
  for (int l = 0; l < 32; l++)
    if (simt::lane_alive(w, l) && (uint32_t)b->val[w * 32 + l] > r) r = (uint32_t)b->val[w * 32 + l];
  simt::warp_barrier();
  return r;
}
static inline int __any_sync(uint32_t m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline void __syncwarp(uint32_t = 0xffffffffu) { simt::warp_barrier(); }
static inline void __syncthreads() { simt::cta_barrier(); }
static inline int __syncthreads_or(int pred) {
  simt::Block *b = simt::g_block;
  b->vote[b->cur] = pred ? 1u : 0u;
  simt::cta_barrier();
  int r = 0;
  for (int t = 0; t < b->n; t++)
    if (!b->done[(size_t)t] && b->vote[t]) r = 1;
  simt::cta_barrier();
  return r;
}
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMin(unsigned *p, unsigned v) { unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
static inline unsigned atomicMax(unsigned *p, unsigned v) { unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { unsigned long long o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { unsigned long long o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned atomicAdd_system(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline uint16_t min(uint16_t a, uint16_t b) { return a < b ? a : b; }
static inline uint16_t max(uint16_t a, uint16_t b) { return a > b ? a : b; }
static inline uint8_t min(uint8_t a, uint8_t b) { return a < b ? a : b; }
static inline uint8_t max(uint8_t a, uint8_t b) { return a > b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

namespace simt {
template <class G>
inline Dim3 to_dim3(const G &g) { return Dim3((unsigned)g.x, (unsigned)g.y, (unsigned)g.z); }
inline Dim3 to_dim3(int g) { return Dim3((unsigned)g); }
inline Dim3 to_dim3(unsigned g) { return Dim3(g); }
extern int g_os_threads;  // OS threads that run blocks side by side (set by the harness)

// SVO_LAUNCH(grid, block, stream, kernel)(args...) in the emulated build
template <class G, class B, class... P>
inline auto launcher(const G &grid, const B &block, void (*kernel)(P...)) {
  const Dim3 g = to_dim3(grid), b = to_dim3(block);
  return [g, b, kernel](P... args) { run_grid(g, b, [&] { kernel(args...); }, g_os_threads); };
}
}  // namespace simt
