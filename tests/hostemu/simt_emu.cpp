// simt_emu.cpp -- TEST INFRASTRUCTURE: scheduler of the coroutine SIMT emulator (see simt_emu.h).
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
  // round-robin; a full round in which nobody exits and no collective completes means the block is stuck
  unsigned long stuck_rounds = 0;
  while (b.alive > 0) {
    const int alive0 = b.alive;
    unsigned gens0 = b.cta_bar.gen;
    for (int w = 0; w * 32 < nthreads; w++) gens0 += b.warp_bar[w].gen;
    for (int t = 0; t < nthreads; t++) {
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
