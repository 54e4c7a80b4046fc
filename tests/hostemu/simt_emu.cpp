// simt_emu.cpp -- TEST INFRASTRUCTURE: scheduler of the coroutine SIMT emulator (see simt_emu.h).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>

#include "simt_emu.h"
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace simt {

thread_local Block *g_block = nullptr;
thread_local Idx3 g_threadIdx, g_blockIdx;
thread_local Dim3 g_blockDim, g_gridDim;
int g_os_threads = 1;

static void trampoline() {
  Block *b = g_block;
  b->body();
  const int t = b->cur;
  b->done[(size_t)t] = 1;
  b->alive--;
  b->warp_alive[t >> 5]--;
  swapcontext(&b->ctx[(size_t)t], &b->main);
}

static void run_block(Block &b, int nthreads, const std::function<void()> &kernel) {
  b.n = b.alive = nthreads;
  b.body = kernel;
  b.ctx.resize((size_t)nthreads);
  b.done.assign((size_t)nthreads, 0);
  while ((int)b.stacks.size() < nthreads) b.stacks.push_back((char *)malloc(Block::kStack));
  for (int w = 0; w < Block::kMaxThreads / 32; w++) {
    const int left = nthreads - w * 32;
    b.warp_alive[w] = left >= 32 ? 32 : (left > 0 ? left : 0);
    b.warp_bar[w] = Bar();
  }
  b.cta_bar = Bar();
  for (int t = 0; t < nthreads; t++) {
    getcontext(&b.ctx[(size_t)t]);
    b.ctx[(size_t)t].uc_stack.ss_sp = b.stacks[(size_t)t];
    b.ctx[(size_t)t].uc_stack.ss_size = Block::kStack;
    b.ctx[(size_t)t].uc_link = nullptr;
    makecontext(&b.ctx[(size_t)t], trampoline, 0);
  }
  g_block = &b;
  // One round resumes every live thread once, each running up to its next collective: in thread order, or -- with
  // SVO_EMU_SHUFFLE=<seed> in the environment -- in a fresh random order every round, which makes a kernel that reads what
  // another lane wrote without a barrier in between fail instead of passing by luck of the order.
  // A full round in which nobody exits and no collective completes means the block is stuck.
  static const char *shuffle_env = getenv("SVO_EMU_SHUFFLE");
  uint64_t rng = shuffle_env ? (uint64_t)strtoull(shuffle_env, nullptr, 10) * 0x9E3779B97F4A7C15ull + g_blockIdx.x * 1315423911ull + g_blockIdx.y + 1 : 0;
  std::vector<int> order((size_t)nthreads);
  for (int t = 0; t < nthreads; t++) order[(size_t)t] = t;
  unsigned long stuck_rounds = 0;
  while (b.alive > 0) {
    if (shuffle_env)
      for (int t = nthreads - 1; t > 0; t--) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        std::swap(order[(size_t)t], order[(size_t)((rng >> 33) % (uint64_t)(t + 1))]);
      }
    const int alive0 = b.alive;
    unsigned gens0 = b.cta_bar.gen;
    for (int w = 0; w * 32 < nthreads; w++) gens0 += b.warp_bar[w].gen;
    for (int k = 0; k < nthreads; k++) {
      const int t = order[(size_t)k];
      if (b.done[(size_t)t]) continue;
      b.cur = t;
      g_threadIdx.x = (unsigned)t;
      swapcontext(&b.main, &b.ctx[(size_t)t]);
    }
    unsigned gens1 = b.cta_bar.gen;
    for (int w = 0; w * 32 < nthreads; w++) gens1 += b.warp_bar[w].gen;
    if (b.alive == alive0 && gens1 == gens0) {
      if (++stuck_rounds > 4) {
        fprintf(stderr, "simt_emu: block (%u,%u) deadlocked: %d threads wait in a collective nobody else reaches\n", g_blockIdx.x,
                g_blockIdx.y, b.alive);
        abort();
      }
    } else {
      stuck_rounds = 0;
    }
  }
  g_block = nullptr;
}

void run_grid(Dim3 grid, Dim3 block, const std::function<void()> &kernel, int os_threads) {
  const unsigned long nblocks = (unsigned long)grid.x * grid.y * grid.z;
  const int nthreads = (int)(block.x * block.y * block.z);
  std::atomic<unsigned long> next(0);
  auto worker = [&] {
    Block b;
    g_blockDim = block;
    g_gridDim = grid;
    for (;;) {
      const unsigned long i = next.fetch_add(1);
      if (i >= nblocks) break;
      g_blockIdx.x = (unsigned)(i % grid.x);
      g_blockIdx.y = (unsigned)((i / grid.x) % grid.y);
      g_blockIdx.z = (unsigned)(i / ((unsigned long)grid.x * grid.y));
      g_threadIdx = Idx3();
      run_block(b, nthreads, kernel);
    }
    for (char *s : b.stacks) free(s);
  };
  if (os_threads <= 1) { worker(); return; }
  std::vector<std::thread> th;
  for (int i = 0; i < os_threads; i++) th.emplace_back(worker);
  for (auto &t : th) t.join();
}

}  // namespace simt

// ---- self-test of the emulator's collective semantics (called from tests/test_hostemu.py) -------------------------
#include "cuda_host_shim.h"
#define threadIdx (simt::g_threadIdx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
namespace {
void k_selftest(unsigned *out) {
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  unsigned *o = out + (size_t)blockIdx.x * 8u * blockDim.x + (size_t)tid * 8u;
  // every third thread leaves at once: collectives must complete among the rest (sm_70+ semantics)
  if (tid % 3u == 2u) { o[0] = 0xDEADu; return; }
  o[0] = __ballot_sync(0xffffffffu, (lane & 1u) != 0u);        // exited lanes vote 0
  o[1] = __shfl_sync(0xffffffffu, tid * 7u + 1u, __ffs(__ballot_sync(0xffffffffu, 1)) - 1);  // from the lowest live lane
  o[2] = __reduce_add_sync(0xffffffffu, lane);
  o[3] = (unsigned)__syncthreads_or(tid == 64u);                // one live thread of warp 2 says yes
  __shared__ unsigned s_sum[4];
  if (tid == 0u) s_sum[0] = s_sum[1] = s_sum[2] = s_sum[3] = 0u;
  __syncthreads();
  atomicAdd(&s_sum[warp], 1u);
  __syncthreads();
  o[4] = s_sum[warp];
  // a loop whose trip count differs per lane, with a vote in every trip (the kernels' refill pattern)
  unsigned trips = 0, left = lane % 5u;
  for (;;) {
    const bool busy = left > 0u;
    if (busy) left--;
    trips++;
    if (__ballot_sync(0xffffffffu, left > 0u) == 0u) break;
  }
  o[5] = trips;
}
}  // namespace

extern "C" int emu_selftest(unsigned *out, int blocks, int os_threads) {
  simt::run_grid(simt::Dim3((unsigned)blocks), simt::Dim3(128), [&] { k_selftest(out); }, os_threads);
  return 0;
}
