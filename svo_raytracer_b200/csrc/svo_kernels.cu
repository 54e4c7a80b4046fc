// svo_kernels.cu -- __global__ entry points of the SVO trace path (sm_100a).
// Launch geometry replaces glDispatchCompute(ceil(W/8), ceil(H/8), 1) with
// local size 8x8 (reference src/engine/Main.java:108-109,285;
// src/shaders/svotrace.comp:648).
#include "svo_kernels.h"
#include "svo_trace.cuh"

// Kernel launch.  The CPU test suite compiles this file with g++ and runs the __global__ functions on a coroutine
// SIMT emulator (tests/hostemu/simt_emu.h, SVO_HOST_EMU): same kernels, same launch geometry, no GPU.
#ifdef SVO_HOST_EMU
#define SVO_LAUNCH(grid, block, stream, ...) simt::launcher(grid, block, __VA_ARGS__)
#else
#define SVO_LAUNCH(grid, block, stream, ...) __VA_ARGS__<<<grid, block, 0, stream>>>
#endif

namespace svo {

// ---------------------------------------------------------------------------
// Kernel variant 0: one thread per pixel.  A warp owns an 8x4 pixel tile (so
// its 32 primary rays share the upper octree levels and its stores fill whole
// 32-byte sectors); a 128-thread CTA owns 16x8 pixels.
// ---------------------------------------------------------------------------
// (CTAs are launched top to bottom: a scrambled row order was measured 5-15 % slower -- CTAs of neighbouring rows
// running together share their octree working set in L1/L2, which outweighs the better tail balance.)
template <bool FAST, bool AUX, bool BOX, int STACK = 0>  // STACK 1: 16-byte stack entries (kernel variant 14: variant 10 with band interleaving)
__global__ void __launch_bounds__(128, 8) k_render_tile(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1,
                                                     int band_stride, int band_offset, int band_ctas) {
  // A band is band_ctas consecutive CTA rows (8 image rows each).  This launch renders bands number
  // i * band_stride + band_offset (counted from y0): with stride = number of GPUs and offset = rank this is the
  // interleaved image partition of the multi-GPU mode; stride 1 is the whole image.
  // CTA = 128 threads: 16x8 pixels (2x2 warps); CTA = 64 threads: 8x8 pixels (1x2 warps)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_x = blockDim.x >> 6;
  const int x = (blockIdx.x * warps_x + (warp % warps_x)) * 8 + (lane & 7);
  const int band = (int)blockIdx.y / band_ctas, in_band = (int)blockIdx.y % band_ctas;
  const int y = y0 + ((band * band_stride + band_offset) * band_ctas + in_band) * 8 + (warp / warps_x) * 4 + (lane >> 3);
  if (x >= W || y >= y1) return;
  shade_pixel<FAST, AUX, false, BOX, false, STACK>(sc, f, pl, W, H, x, y);
}

// ---------------------------------------------------------------------------
// Kernel variant 1: persistent threads with warp-level ray fetch (the
// "while-while" organisation of Aila & Laine 2009, applied per cast).  One CTA
// per SM slot loops until the frame's tile queue is empty.  Every lane is a
// small state machine -- fetch a pixel, set up a cast, traverse, shade, maybe
// cast again -- and the warp leaves the traversal loop as soon as kRefill
// lanes have finished their cast, so those lanes are re-armed (next cast of the
// same pixel, or a new pixel from the warp's current 8x4 tile) instead of
// idling until the slowest ray of the warp terminates.  Ray iteration counts
// range from 1 to 1500 within a warp (bounce rays especially), which is where
// variant 0 loses most of its issue slots.
// ---------------------------------------------------------------------------
constexpr int kRefill = 8;  // leave the traversal loop when this many lanes wait for work

template <bool FAST, bool AUX>
__global__ void __launch_bounds__(128) k_render_persistent(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1,
                                                           unsigned int *__restrict__ tile_counter) {
  enum { NEED_PIXEL = 0, NEED_SETUP = 1, TRAVERSING = 2, CAST_DONE = 3, EXHAUSTED = 4 };
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int tiles_x = (W + 7) >> 3, tiles_y = (y1 - y0 + 3) >> 2;
  const unsigned ntiles = (unsigned)tiles_x * (unsigned)tiles_y;

  uint2 stk[kMaxScale + 1];
  Trav<FAST> T;
  Pixel P;
  int state = NEED_PIXEL, status = TRAV_CONTINUE;
  unsigned pool_next = 0, pool_end = 0;  // the warp's current tile: pixel slots [pool_next, pool_end)
  bool queue_empty = false;

  for (;;) {
    // ---- hand pixels to idle lanes ---------------------------------------------
    unsigned want = __ballot_sync(0xffffffffu, state == NEED_PIXEL);
    while (want != 0u) {
      if (pool_next == pool_end) {
        if (queue_empty) break;
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntiles) { queue_empty = true; break; }
        pool_next = t * 32u;
        pool_end = pool_next + 32u;
      }
      const unsigned avail = pool_end - pool_next;
      const unsigned rank = __popc(want & lt_mask);
      const bool mine = ((want >> lane) & 1u) != 0u && rank < avail;
      if (mine) {
        const unsigned slot = pool_next + rank, tile = slot >> 5, k = slot & 31u;
        const int x = (int)(tile % (unsigned)tiles_x) * 8 + (int)(k & 7u);
        const int y = y0 + (int)(tile / (unsigned)tiles_x) * 4 + (int)(k >> 3);
        if (x < W && y < y1) {
          if (pixel_begin(f, pl, W, H, x, y, P)) state = NEED_SETUP;
          else pixel_store<AUX>(sc, f, pl, W, P);  // no cast wanted (mode 4): done, lane stays idle
        }
      }
      pool_next += min(avail, (unsigned)__popc(want));
      want = __ballot_sync(0xffffffffu, state == NEED_PIXEL);  // unserved lanes, and lanes whose slot was outside the image
    }
    if (state == NEED_PIXEL && queue_empty && pool_next == pool_end) state = EXHAUSTED;
    if (state == NEED_SETUP) {
      T.setup(sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, nullptr);
      state = TRAVERSING;
    }
    if (__ballot_sync(0xffffffffu, state == TRAVERSING) == 0u) {
      if (__ballot_sync(0xffffffffu, state != EXHAUSTED) == 0u) break;  // frame done for this warp
      continue;
    }
    // ---- traverse until enough lanes have finished their cast ----------------------
    const int busy0 = __popc(__ballot_sync(0xffffffffu, state == TRAVERSING));
    for (;;) {
      if (state == TRAVERSING) {
        status = T.step(sc, stk, nullptr);
        if (status != TRAV_CONTINUE) state = CAST_DONE;
      }
      const int busy = __popc(__ballot_sync(0xffffffffu, state == TRAVERSING));
      if (busy + kRefill <= busy0 || busy == 0) break;
    }
    // ---- shade finished casts; they either cast again or release the lane ----------
    if (state == CAST_DONE) {
      if (pixel_finish_cast(sc, f, P, T.export_hit(status))) state = NEED_SETUP;
      else {
        pixel_store<AUX>(sc, f, pl, W, P);
        state = NEED_PIXEL;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Kernel variant 6: variant 0 made CTA-synchronous between casts, with the rays of every cast after the first
// re-dealt inside the CTA by direction octant.  Primary rays are traced by the thread that owns the pixel (an
// 8x4 tile per warp is already coherent).  Bounce / shadow rays of the CTA's 128 pixels are counting-sorted in
// shared memory by octant_mask (svotrace.comp:238-241; finished pixels last), traced by whichever thread gets
// them, and the 28-byte end state of the cast (HitState) is handed back to the owning thread, which runs the
// code after the loop and the shading exactly as in variant 0.  Rays of one octant visit children in the same
// order and -- starting from neighbouring pixels -- share their descent, so a warp agrees on PUSH/ADVANCE/POP far
// more often; finished pixels collect in whole warps that retire at once.  Per-ray arithmetic is untouched.
// ---------------------------------------------------------------------------
template <bool FAST, bool AUX, bool BOX>
__global__ void __launch_bounds__(128, 8) k_render_tile_binned(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1) {
  __shared__ float s_ray[7][128];        // origin xyz, dir xyz, cone flag
  __shared__ unsigned char s_owner[128];  // slot -> thread that owns the pixel
  __shared__ uint32_t s_hit[7][128];      // HitState per owning thread
  __shared__ unsigned short s_cnt[4][10]; // per warp, per key
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const int y = y0 + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
  const bool valid = x < W && y < y1;
  Pixel P;
  bool wants = valid && pixel_begin(f, pl, W, H, x, y, P);
  uint2 stk[kMaxScale + 1];
  for (int cast = 0; __syncthreads_or(wants); cast++) {
    HitState hs;
    if (cast == 0) {
      if (wants) {
        Trav<FAST, false, BOX> T;
        T.setup(sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, nullptr);
        hs = T.export_hit((T.outside_box() || T.nan_ray(nullptr)) ? TRAV_MISS : T.run(sc, stk, nullptr));
      }
    } else {
      // ---- counting sort of the CTA's rays by octant (key 8 = no ray) ----
      const unsigned key = wants ? ((P.dir.x > 0.0f ? 1u : 0u) | (P.dir.y > 0.0f ? 2u : 0u) | (P.dir.z > 0.0f ? 4u : 0u)) : 8u;
      unsigned my_rank = 0;
#pragma unroll
      for (unsigned b = 0; b < 9; b++) {
        const unsigned m = __ballot_sync(0xffffffffu, key == b);
        if (key == b) my_rank = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[warp][b] = (unsigned short)__popc(m);
      }
      __syncthreads();
      unsigned base = 0;
      for (unsigned b = 0; b < 9; b++)
        for (int w = 0; w < 4; w++)
          if (b < key || (b == key && w < warp)) base += s_cnt[w][b];
      const unsigned slot = base + my_rank;
      s_owner[slot] = (unsigned char)tid;
      s_ray[0][slot] = P.origin.x; s_ray[1][slot] = P.origin.y; s_ray[2][slot] = P.origin.z;
      s_ray[3][slot] = P.dir.x; s_ray[4][slot] = P.dir.y; s_ray[5][slot] = P.dir.z;
      s_ray[6][slot] = (wants && P.cone) ? 1.0f : 0.0f;
      unsigned nrays = 0;
      for (int w = 0; w < 4; w++)
        for (unsigned b = 0; b < 8; b++) nrays += s_cnt[w][b];
      __syncthreads();
      // ---- trace the ray in slot `tid`, hand its end state to the owner ----
      if ((unsigned)tid < nrays) {
        Trav<FAST, false, BOX> T;
        T.setup(sc, mk3(s_ray[0][tid], s_ray[1][tid], s_ray[2][tid]), mk3(s_ray[3][tid], s_ray[4][tid], s_ray[5][tid]), f.maxDepth,
                s_ray[6][tid] != 0.0f, f.coneDepth, nullptr);
        const HitState r = T.export_hit((T.outside_box() || T.nan_ray(nullptr)) ? TRAV_MISS : T.run(sc, stk, nullptr));
        const unsigned o = s_owner[tid];
        s_hit[0][o] = r.pidx; s_hit[1][o] = r.meta; s_hit[2][o] = r.ipx; s_hit[3][o] = r.ipy; s_hit[4][o] = r.ipz;
        s_hit[5][o] = __float_as_uint(r.t_min); s_hit[6][o] = r.iter;
      }
      __syncthreads();
      if (wants) {
        hs.pidx = s_hit[0][tid]; hs.meta = s_hit[1][tid]; hs.ipx = s_hit[2][tid]; hs.ipy = s_hit[3][tid]; hs.ipz = s_hit[4][tid];
        hs.t_min = __uint_as_float(s_hit[5][tid]); hs.iter = s_hit[6][tid];
      }
    }
    if (wants) wants = pixel_finish_cast(sc, f, P, hs);
  }
  if (valid) pixel_store<AUX>(sc, f, pl, W, P);
}

// ---------------------------------------------------------------------------
// Kernel variants 9..12: variant 0 with one thing changed each (separately rounded arithmetic only).
//   9  SmemStack: the parent stack in shared memory (14 levels x 128 threads x 8 B = 14 KB per CTA) instead of local
//      memory -- stack accesses cannot miss and stop competing with descriptors for L1; frames with maxDepth <= 14
//  10  WideStack: 16-byte stack entries carrying the parent's descriptor -- a POP is one LDL.128 instead of LDL.64
//      followed by a dependent descriptor load
//  11 / 12  __launch_bounds__(128, 7) / (128, 6): 72 / 80 registers per thread instead of 64
//  13  variant 10 with the pipe-balanced loop body (Trav BAL): integer adds, moves, select chains and constant shifts
//      of the loop as IMADs on the FMA pipe (ncu: ALU pipe 75 % busy, math-pipe throttle stalls, FMA pipe 20 %).
//      Static count of the loop: ALU-pipe instructions 64 -> 44, FMA-pipe 26 -> 43, 6 more constant loads.
// ---------------------------------------------------------------------------
template <bool AUX, bool BOX, int STACK, int BAL = 0>
__device__ __forceinline__ void tile_body(const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H, int y0, int y1) {
  __shared__ uint2 s_stack[STACK == 2 ? kSmemStackLevels : 1][kSmemStackStride];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const int y = y0 + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
  if (x >= W || y >= y1) return;
  shade_pixel<false, AUX, false, BOX, false, STACK, BAL>(sc, f, pl, W, H, x, y, nullptr, &s_stack[0][threadIdx.x]);
}
template <bool AUX, bool BOX, int BAL = 15, int MINB = 8>  // BAL: which parts of the loop run as IMADs (mask, see Trav); 15 = variant 13
__global__ void __launch_bounds__(128, MINB) k_render_tile_balanced(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1) {
  tile_body<AUX, BOX, 1, BAL>(sc, f, pl, W, H, y0, y1);  // MINB 7 / 6: 72 / 80 registers, so that the loop's constants (sc.one, ...) need not be re-loaded every iteration
}
template <bool AUX, bool BOX, int STACK>
__global__ void __launch_bounds__(128, 8) k_render_tile_stack(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1) {
  tile_body<AUX, BOX, STACK>(sc, f, pl, W, H, y0, y1);
}
template <bool AUX, bool BOX, int MINB>
__global__ void __launch_bounds__(128, MINB) k_render_tile_regs(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1) {
  tile_body<AUX, BOX, 0>(sc, f, pl, W, H, y0, y1);
}

// ---------------------------------------------------------------------------
// Kernel variant 17: variant 10 as persistent warps that take whole 8x4-pixel tiles from a queue (one atomic per tile,
// warp-level fetch).  Nothing about the per-pixel work changes -- no lane is re-armed inside a tile, which is what cost
// variants 1 / 2 / 7 / 8 their coherence -- only WHO renders a tile does: a warp that finishes early takes the next tile at
// once instead of idling until the three other warps of its CTA are done (ncu, variant 10: achieved occupancy 42-44 % of
// the 50 % the registers allow), and a launch has no wave quantisation: 2 025 CTAs over 1 184 CTA slots = 1.7 waves is what
// bounds one GPU's share of a 1080p frame in the 8-GPU tile partition.  Tiles are numbered so that four consecutive
// indices form a 16x8 block and blocks run left to right, top to bottom: warps that fetch together work on neighbouring
// pixels, like the CTAs of variant 10 (L1 / L2 locality).  band_stride / band_offset select the interleaved bands of one
// GPU as in k_render_tile.  The queue head is reset by the last CTA to leave, so a frame is one launch and nothing else;
// the same CTA then signals the frame-complete fences of the multi-GPU partition (FenceList; n = 0: none), which saves
// the separate signal kernel and its launch latency.
// ---------------------------------------------------------------------------
template <bool AUX, bool BOX>
__global__ void __launch_bounds__(128, 8) k_render_tile_queue(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1,
                                                           int band_stride, int band_offset, int band_ctas, unsigned int nblocks_y,
                                                           unsigned int *__restrict__ queue, FenceList fl) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned blocks_x = ((unsigned)W + 15u) >> 4;
  const unsigned ntiles = blocks_x * nblocks_y * 4u;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(queue, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= ntiles) break;
    const unsigned b = t >> 2, sub = t & 3u;
    const unsigned bx = b % blocks_x, by = b / blocks_x;
    const int band = (int)by / band_ctas, in_band = (int)by % band_ctas;
    const int x = (int)(bx * 16u + (sub & 1u) * 8u + (lane & 7u));
    const int y = y0 + ((band * band_stride + band_offset) * band_ctas + in_band) * 8 + (int)(sub >> 1) * 4 + (int)(lane >> 3);
    if (x < W && y < y1) shade_pixel<false, AUX, false, BOX, false, 1, 15>(sc, f, pl, W, H, x, y);
  }
  // last CTA out: reset the queue for the next launch, then tell the frame's owner that this GPU's pixels are stored
  __syncthreads();
  if (threadIdx.x == 0) {
    if (fl.n > 0) __threadfence_system();  // this CTA's pixel stores (possibly into a peer's memory) are ordered before its ticket
    else __threadfence();
    const unsigned ticket = atomicAdd(queue + 1, 1u);
    if (ticket == gridDim.x - 1u) {
      queue[0] = 0u;
      queue[1] = 0u;
      if (fl.n > 0) {
        __threadfence_system();  // every CTA's stores (ordered before its ticket) are visible system-wide before the bumps
        for (int i = 0; i < fl.n; i++) atomicAdd_system(fl.p[i], 1u);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Kernel variants 7 / 8: tile kernel for the primary rays + warp-local lane refill for every later cast.
// The divergence model (tools/simt_model.py) puts the bounce casts of the bench workload at ~29 % lane utilisation,
// mostly trip-count divergence: a warp waits for its longest ray.  Here a warp owns PIX 8x4 tiles (PIX pixels per
// thread).  Primary rays run exactly as in variant 0, one tile after the other (they are coherent).  The rays of
// every later cast go into a queue in the warp's own shared memory; lanes take rays from it, and as soon as
// kRefillIdle lanes have finished their ray the warp leaves the traversal loop, parks the end states (HitState, 28
// bytes, in the ray's own queue slot) and hands those lanes the next rays.  When the queue is dry every thread
// picks up the end states of its own pixels and shades them at full width.  Unlike variant 1 nothing but
// Trav::setup runs at partial lane utilisation, and unlike variant 6 nothing synchronises wider than a warp.
// ---------------------------------------------------------------------------
constexpr int kRefillIdle = 8;   // lanes that must be free before the warp leaves the traversal loop for a refill
constexpr int kRefillBatch = 8;  // loop iterations between two votes

template <bool FAST, bool AUX, bool BOX, int PIX>
__global__ void __launch_bounds__(128, 8) k_render_tile_refill(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1) {
  constexpr int TX = 2, TY = PIX / 2;        // tiles of one warp: 2 x TY (PIX = 2: 16x4 pixels, PIX = 4: 16x8 pixels)
  constexpr int SLOTS = PIX * 32;
  __shared__ uint32_t s_q[4][7][SLOTS];      // per warp: ray (origin, dir, cone flag) on the way in, HitState on the way out
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t(*q)[SLOTS] = s_q[warp];
  const int wx = (blockIdx.x * 2 + (warp & 1)) * (TX * 8), wy = y0 + (blockIdx.y * 2 + (warp >> 1)) * (TY * 4);

  // Pixel states live in local memory between phases (indexed dynamically: the code of a phase exists once, not PIX
  // times -- the unrolled kernel was 13 k instructions); the pixel being worked on is in registers.
  Pixel P[PIX];
  unsigned valid = 0, wants = 0;  // bit k: pixel k is inside the image / wants another cast
  uint2 stk[kMaxScale + 1];
  // ---- primary casts, tile by tile, as in variant 0 ----
#pragma unroll 1
  for (int k = 0; k < PIX; k++) {
    const int x = wx + (k % TX) * 8 + (int)(lane & 7u), y = wy + (k / TX) * 4 + (int)(lane >> 3);
    if (x < W && y < y1) {
      valid |= 1u << k;
      Pixel cur;
      bool more = pixel_begin(f, pl, W, H, x, y, cur);
      if (more) {
        Trav<FAST, false, BOX> T;
        T.setup(sc, cur.origin, cur.dir, f.maxDepth, cur.cone, f.coneDepth, nullptr);
        const HitState hs = T.export_hit((T.outside_box() || T.nan_ray(nullptr)) ? TRAV_MISS : T.run(sc, stk, nullptr));
        more = pixel_finish_cast(sc, f, cur, hs);
      }
      if (more) wants |= 1u << k;
      P[k] = cur;
    }
  }
  // ---- later casts: one round per cast index, rays of the warp's PIX tiles in one queue ----
  for (;;) {
    unsigned count = 0;
    unsigned slots = 0;  // byte k: queue slot of pixel k's ray
#pragma unroll 1
    for (int k = 0; k < PIX; k++) {
      const bool w = ((wants >> k) & 1u) != 0u;
      const unsigned m = __ballot_sync(0xffffffffu, w);
      const unsigned s_ = count + __popc(m & lt_mask);
      count += __popc(m);
      if (w) {
        slots |= s_ << (8 * k);
        q[0][s_] = __float_as_uint(P[k].origin.x); q[1][s_] = __float_as_uint(P[k].origin.y); q[2][s_] = __float_as_uint(P[k].origin.z);
        q[3][s_] = __float_as_uint(P[k].dir.x); q[4][s_] = __float_as_uint(P[k].dir.y); q[5][s_] = __float_as_uint(P[k].dir.z);
        q[6][s_] = P[k].cone ? 1u : 0u;
      }
    }
    if (count == 0) break;
    __syncwarp();
    unsigned head = 0, mine = 0;
    int busy = 0;
    Trav<FAST, false, BOX> T;
    for (;;) {
      // hand rays to idle lanes (a ray that ends before the loop -- outside the content box, NaN -- frees its lane at once)
      unsigned idle = __ballot_sync(0xffffffffu, !busy);
      while (idle != 0u && head < count) {
        const unsigned rank = __popc(idle & lt_mask);
        if (!busy && head + rank < count) {
          mine = head + rank;
          T.setup(sc, mk3(__uint_as_float(q[0][mine]), __uint_as_float(q[1][mine]), __uint_as_float(q[2][mine])),
                  mk3(__uint_as_float(q[3][mine]), __uint_as_float(q[4][mine]), __uint_as_float(q[5][mine])), f.maxDepth, q[6][mine] != 0u,
                  f.coneDepth, nullptr);
          if (T.outside_box() || T.nan_ray(nullptr)) {
            const HitState r = T.export_hit(TRAV_MISS);
            q[0][mine] = r.pidx; q[1][mine] = r.meta; q[2][mine] = r.ipx; q[3][mine] = r.ipy; q[4][mine] = r.ipz;
            q[5][mine] = __float_as_uint(r.t_min); q[6][mine] = r.iter;
          } else {
            busy = 1;
          }
        }
        head += min(count - head, (unsigned)__popc(idle));
        idle = __ballot_sync(0xffffffffu, !busy);
      }
      const int busy0 = __popc(__ballot_sync(0xffffffffu, busy));
      if (busy0 == 0) break;  // queue dry and every ray finished
      // traverse until enough lanes are free to make a refill worth its cost (or, with the queue dry, until all are
      // done).  The warp votes once per kRefillBatch iterations: a vote per iteration costs 13 instructions on top of
      // the ~100 of the iteration, a lane that finishes inside a batch idles for < kRefillBatch iterations.
      const int limit = head < count ? max(busy0 - kRefillIdle, 0) : 0;
      for (;;) {
        if (busy) {
          int status = TRAV_CONTINUE;
#pragma unroll 1
          for (int i = 0; i < kRefillBatch && status == TRAV_CONTINUE; i++) status = T.step(sc, stk, nullptr);
          if (status != TRAV_CONTINUE) {
            busy = 0;
            const HitState r = T.export_hit(status);
            q[0][mine] = r.pidx; q[1][mine] = r.meta; q[2][mine] = r.ipx; q[3][mine] = r.ipy; q[4][mine] = r.ipz;
            q[5][mine] = __float_as_uint(r.t_min); q[6][mine] = r.iter;
          }
        }
        if (__popc(__ballot_sync(0xffffffffu, busy)) <= limit) break;
      }
    }
    __syncwarp();
    // ---- every thread shades its own pixels from the parked end states ----
#pragma unroll 1
    for (int k = 0; k < PIX; k++) {
      if ((wants >> k) & 1u) {
        const unsigned s_ = (slots >> (8 * k)) & 0xFFu;
        HitState hs;
        hs.pidx = q[0][s_]; hs.meta = q[1][s_]; hs.ipx = q[2][s_]; hs.ipy = q[3][s_]; hs.ipz = q[4][s_];
        hs.t_min = __uint_as_float(q[5][s_]); hs.iter = q[6][s_];
        Pixel cur = P[k];
        if (!pixel_finish_cast(sc, f, cur, hs)) wants &= ~(1u << k);
        P[k] = cur;
      }
    }
    __syncwarp();  // the slots are rewritten by the next round
  }
#pragma unroll 1
  for (int k = 0; k < PIX; k++)
    if ((valid >> k) & 1u) {
      Pixel cur = P[k];
      pixel_store<AUX>(sc, f, pl, W, cur);
    }
}

// ---------------------------------------------------------------------------
// Kernel variant 15: the last cast of a mode-0 pixel (the diffuse bounce of the metric's configuration) in a kernel of its
// own, with lane refill.  Lesson of variants 7 / 8: refilling pays only if the state that waits for a ray is small and
// not in local memory.  k_split_primary is variant 10 up to the point where a pixel wants its last cast; it then appends
// a 80-byte record -- the ray, and the nine floats + two flags the code after the cast still reads (stale normal,
// accumulated colour, mask, mirror flag, pixel) -- to a global queue (warp-aggregated atomic) instead of tracing.
// k_split_bounce is persistent: a warp takes 128 records, traces them with lanes re-armed when kRefillIdle are free,
// parks each end state (six words: the iteration count is not observable in mode 0) in its record, and when the chunk
// is done finishes its 128 pixels at full width through the same pixel_finish_cast / pixel_store as every other
// kernel, on a Pixel rebuilt from the record (fields the last cast's shading never reads are zero).
// ---------------------------------------------------------------------------
constexpr unsigned kSplitChunk = 128;

// PRE (variant 16): the primary kernel also runs Trav::setup for the deferred cast -- at full width -- and queues its result
// (ten words) instead of origin and direction, so that a refill in k_split_bounce is a dozen loads instead of ~230
// instructions; casts that end before the loop (outside the content box, NaN) are finished here and never queued.
template <bool BOX, bool PRE>
__global__ void __launch_bounds__(128, 8) k_split_primary(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1, SplitQueue q) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 16 + (int)(warp & 1u) * 8 + (int)(lane & 7u);
  const int y = y0 + blockIdx.y * 8 + (int)(warp >> 1) * 4 + (int)(lane >> 3);
  const bool valid = x < W && y < y1;
  Pixel P;
  bool more = valid && pixel_begin(f, pl, W, H, x, y, P);
  bool defer = false;
  uint4 wide[kMaxScale + 1];
  WideStack ws;
  ws.p = wide;
  while (more) {
    if (P.cast_i >= 1 && P.cast_i + 1 >= f.casts) { defer = true; break; }  // the last cast of a path: k_split_bounce traces it
    uint32_t loops = 0;
    const bool attrs = cast_needs_attrs(f, P);
    const bool hit = cast_ray_on<false, false, BOX, false>(ws, sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, P.res, loops, nullptr, attrs);
    if (!attrs && hit && f.renderMode == 0) {
      pixel_after_last_hit(P);
      more = false;
    } else {
      more = pixel_after_cast(f, P, hit, loops);
    }
  }
  uint32_t w[10];
  if (PRE && defer) {
    Trav<false, false, BOX> T;
    T.setup(sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, nullptr);
    if (T.outside_box() || T.nan_ray(nullptr)) {  // ends before the loop: finish the pixel here
      pixel_finish_cast(sc, f, P, T.export_hit(TRAV_MISS));
      defer = false;
    } else {
      T.save_setup(w);
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, defer);
  if (m != 0u) {
    unsigned base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(q.counters + 0, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (defer) {
      const size_t s_ = (size_t)base + (size_t)__popc(m & ((1u << lane) - 1u));
      const bool mirror = f.mirrorValue != 0 && P.res.value == (uint32_t)f.mirrorValue;
      const uint32_t pix = (uint32_t)((size_t)y * (size_t)W + (size_t)x) | (mirror ? 0x80000000u : 0u);
      if (PRE) {
        q.q[0][s_] = make_uint4(w[0], w[1], w[2], pix);
        q.q[1][s_] = make_uint4(w[3], w[4], w[5], w[9]);
        q.q[2][s_] = make_uint4(w[6], w[7], w[8], __float_as_uint(P.res.normal.x));
        q.q[3][s_] = make_uint4(__float_as_uint(P.dir.x), __float_as_uint(P.dir.y), __float_as_uint(P.dir.z), __float_as_uint(P.res.normal.y));
        q.q[4][s_] = make_uint4(__float_as_uint(P.res.normal.z), __float_as_uint(P.acc.x), __float_as_uint(P.acc.y), __float_as_uint(P.acc.z));
        q.q[5][s_] = make_uint4(__float_as_uint(P.mask.x), __float_as_uint(P.mask.y), __float_as_uint(P.mask.z), 0u);
      } else {
        q.q[0][s_] = make_uint4(__float_as_uint(P.origin.x), __float_as_uint(P.origin.y), __float_as_uint(P.origin.z), pix);
        q.q[1][s_] = make_uint4(__float_as_uint(P.dir.x), __float_as_uint(P.dir.y), __float_as_uint(P.dir.z), __float_as_uint(P.res.normal.x));
        q.q[2][s_] = make_uint4(__float_as_uint(P.res.normal.y), __float_as_uint(P.res.normal.z), __float_as_uint(P.acc.x), __float_as_uint(P.acc.y));
        q.q[3][s_] = make_uint4(__float_as_uint(P.acc.z), __float_as_uint(P.mask.x), __float_as_uint(P.mask.y), __float_as_uint(P.mask.z));
      }
    }
  }
  if (valid && !defer) pixel_store<false>(sc, f, pl, W, P);
}

template <bool BOX, bool PRE>
__global__ void __launch_bounds__(128, 8) k_split_bounce(SceneView sc, FrameParams f, Planes pl, int W, SplitQueue q) {
  const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
  const uint64_t n = (uint64_t)q.counters[0];  // written by k_split_primary, earlier in the stream
  uint4 wide[kMaxScale + 1];
  WideStack stk;
  stk.p = wide;
  for (;;) {
    unsigned c = 0;
    if (lane == 0) c = atomicAdd(q.counters + 1, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    const uint64_t a = (uint64_t)c * kSplitChunk;
    if (a >= n) break;
    const uint64_t b = a + kSplitChunk < n ? a + kSplitChunk : n;
    // ---- trace the chunk's rays, re-arming lanes as rays finish ----
    uint64_t head = a, mine = 0;
    int busy = 0;
    Trav<false, false, BOX> T;
    for (;;) {
      unsigned idle = __ballot_sync(0xffffffffu, !busy);
      while (idle != 0u && head < b) {
        const unsigned rank = __popc(idle & lt_mask);
        if (!busy && head + rank < b) {
          mine = head + rank;
          const uint4 r0 = q.q[0][mine], r1 = q.q[1][mine];
          if (PRE) {
            const uint4 r2 = q.q[2][mine];
            const uint32_t w[10] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z, r1.w};
            T.restore_setup(sc, w, f.maxDepth, true, f.coneDepth);
            busy = 1;  // casts that end before the loop were finished by k_split_primary
          } else {
            T.setup(sc, mk3(__uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z)),
                    mk3(__uint_as_float(r1.x), __uint_as_float(r1.y), __uint_as_float(r1.z)), f.maxDepth, true, f.coneDepth, nullptr);
            if (T.outside_box() || T.nan_ray(nullptr)) {
              const HitState r = T.export_hit(TRAV_MISS);
              q.q[0][mine] = make_uint4(r.pidx, r.meta, r.ipx, r0.w);
              q.q[4][mine] = make_uint4(r.ipy, r.ipz, __float_as_uint(r.t_min), 0u);
            } else {
              busy = 1;
            }
          }
        }
        const uint64_t left = b - head;
        const unsigned take = (unsigned)__popc(idle);
        head += left < take ? left : take;
        idle = __ballot_sync(0xffffffffu, !busy);
      }
      const int busy0 = __popc(__ballot_sync(0xffffffffu, busy));
      if (busy0 == 0) break;
      // with the set-up already done a refill is a dozen loads: the divergence model then prefers re-arming at 4 free lanes
      const int limit = head < b ? max(busy0 - (PRE ? kRefillIdle / 2 : kRefillIdle), 0) : 0;
      for (;;) {
        if (busy) {
          int status = TRAV_CONTINUE;
#pragma unroll 1
          for (int k = 0; k < kRefillBatch && status == TRAV_CONTINUE; k++) status = T.step(sc, stk, nullptr);
          if (status != TRAV_CONTINUE) {
            busy = 0;
            const HitState r = T.export_hit(status);
            const uint32_t pix = q.q[0][mine].w;
            q.q[0][mine] = make_uint4(r.pidx, r.meta, r.ipx, pix);
            q.q[PRE ? 1 : 4][mine] = make_uint4(r.ipy, r.ipz, __float_as_uint(r.t_min), 0u);
          }
        }
        if (__popc(__ballot_sync(0xffffffffu, busy)) <= limit) break;
      }
    }
    __syncwarp();
    // ---- finish the chunk's pixels at full width ----
#pragma unroll 1
    for (uint64_t i = a + lane; i < b; i += 32) {
      const uint4 r0 = q.q[0][i], r1 = q.q[1][i], r2 = q.q[2][i], r3 = q.q[3][i], r4 = q.q[4][i];
      const uint32_t pix = r0.w & 0x7FFFFFFFu;
      Pixel P;
      P.x = (int)(pix % (uint32_t)W);
      P.y = (int)(pix / (uint32_t)W);
      P.origin = mk3(0.0f, 0.0f, 0.0f);  // dead after the last cast
      P.cone = true;
      P.cast_i = f.casts - 1;
      cast_res_clear(P.res);
      HitState hs;
      hs.pidx = r0.x; hs.meta = r0.y; hs.ipx = r0.z;
      if (PRE) {
        const uint4 r5 = q.q[5][i];
        P.dir = mk3(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z));
        P.res.normal = mk3(__uint_as_float(r2.w), __uint_as_float(r3.w), __uint_as_float(r4.x));
        P.acc = mk3(__uint_as_float(r4.y), __uint_as_float(r4.z), __uint_as_float(r4.w));
        P.mask = mk3(__uint_as_float(r5.x), __uint_as_float(r5.y), __uint_as_float(r5.z));
        hs.ipy = r1.x; hs.ipz = r1.y; hs.t_min = __uint_as_float(r1.z);
      } else {
        P.dir = mk3(__uint_as_float(r1.x), __uint_as_float(r1.y), __uint_as_float(r1.z));
        P.res.normal = mk3(__uint_as_float(r1.w), __uint_as_float(r2.x), __uint_as_float(r2.y));
        P.acc = mk3(__uint_as_float(r2.z), __uint_as_float(r2.w), __uint_as_float(r3.x));
        P.mask = mk3(__uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w));
        hs.ipy = r4.x; hs.ipz = r4.y; hs.t_min = __uint_as_float(r4.z);
      }
      P.res.value = (r0.w & 0x80000000u) ? (uint32_t)f.mirrorValue : ~(uint32_t)f.mirrorValue;  // only `== mirrorValue` is read
      P.color = mk3(0.0f, 0.0f, 0.0f);
      P.depth = 0.0f;
      P.beamDist = 0.0f;
      P.t_floor = 0.0f;
      P.hit_id = kNoHit;
      P.iter = 0;
      P.primary_t = 0.0f;
      hs.iter = 0;  // not observable in render mode 0 without the validation planes
      pixel_finish_cast(sc, f, P, hs);
      pixel_store<false>(sc, f, pl, W, P);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Kernel variant 4 (experiment): variant 0 with the upper octree levels staged in shared memory.  Every CTA
// copies the first kTopDescs descriptors (breadth-first array: levels 0..3 and part of 4) before tracing.
// ---------------------------------------------------------------------------
constexpr unsigned kTopDescs = 512;
__global__ void __launch_bounds__(128, 8) k_render_tile_smem(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1) {
  __shared__ uint2 top[kTopDescs];
  const unsigned ntop = min(sc.ndesc, kTopDescs);
  for (unsigned i = threadIdx.x; i < ntop; i += blockDim.x) top[i] = __ldg(sc.desc + i);
  __syncthreads();
  sc.top = top;
  sc.ntop = ntop;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const int y = y0 + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
  if (x >= W || y >= y1) return;
  shade_pixel<false, false, false, true, true>(sc, f, pl, W, H, x, y);
}

// ---------------------------------------------------------------------------
// Frame-complete fence for the multi-GPU tile partition.  Every GPU, after the kernel that stored its bands into
// the frame owner's planes (peer-mapped over NVLink), bumps a counter in the owner's memory; the owner's stream
// waits until all GPUs have checked in.  Replaces a per-frame NCCL collective (~40 us) by one remote atomic.
// ---------------------------------------------------------------------------
__global__ void k_fence_signal(FenceList fl) {
  __threadfence_system();  // the stores of the preceding kernels in this stream are complete; order the bumps after them
  if ((int)threadIdx.x < fl.n) atomicAdd_system(fl.p[threadIdx.x], 1u);
}
__global__ void k_fence_wait(volatile unsigned int *fence, volatile unsigned int *dead, unsigned int target) {
  const long long t0 = clock64();
  // modular comparison (the counter runs forever).  Give up after ~4e9 cycles (~2 s) instead of hanging the GPU,
  // and latch the failure so that every later wait returns at once.
  while ((int)(*fence - target) < 0) {
    if (*dead == 0xDEADu) break;
    if (clock64() - t0 > 4000000000ll) { *dead = 0xDEADu; break; }
    __nanosleep(100);
  }
  __threadfence_system();
}

// wait, then signal: "frame complete" -> "frame consumed" on the owner without a second launch (nothing reads the frame in between)
__global__ void k_fence_wait_signal(volatile unsigned int *fence, volatile unsigned int *dead, unsigned int target, FenceList fl) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while ((int)(*fence - target) < 0) {
      if (*dead == 0xDEADu) break;
      if (clock64() - t0 > 4000000000ll) { *dead = 0xDEADu; break; }
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
  if ((int)threadIdx.x < fl.n) atomicAdd_system(fl.p[threadIdx.x], 1u);
}

// Instrumented build of variant 0: same traversal, plus the oracle's counters
// (casts, loop iterations, bytes of the reference-layout records the reference
// would have fetched).  bench.py runs it once, outside the timed region, to get
// the algorithmic bytes of the workload; tests compare it with the oracle.
template <bool EXECUTED>  // false: the oracle's counters; true: what the production kernel executes (content box on, early exits)
__global__ void __launch_bounds__(128) k_render_stats(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1,
                                                      unsigned long long *counters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const int y = y0 + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
  RayStats rs;
  rs.casts = rs.iters = rs.record_bytes = 0u;
  if (x < W && y < y1) {
    if (EXECUTED) shade_pixel<false, false, 2, true, false, 1>(sc, f, pl, W, H, x, y, &rs);
    else shade_pixel<false, true, 1>(sc, f, pl, W, H, x, y, &rs);
  }
  const uint32_t c = __reduce_add_sync(0xffffffffu, rs.casts);
  const uint32_t i = __reduce_add_sync(0xffffffffu, rs.iters);
  const uint32_t b = __reduce_add_sync(0xffffffffu, rs.record_bytes);
  if (lane == 0) {
    atomicAdd(counters + 0, (unsigned long long)c);
    atomicAdd(counters + 1, (unsigned long long)i);
    atomicAdd(counters + 2, (unsigned long long)b);
  }
}

// Random-sector gather microbenchmark: the "gather roofline" S of SURVEY 8d.
// Every thread issues `loads` independent 8-byte read-only loads (8 in flight)
// at hashed offsets of a `words`-long uint2 array, i.e. one random 32-byte
// sector each -- the access pattern of a PUSH in cast_ray.
__global__ void __launch_bounds__(256) k_gather_probe(const uint2 *__restrict__ buf, uint64_t words, int loads, uint32_t *__restrict__ sink) {
  uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  uint32_t acc = 0;
  for (int i = 0; i < loads; i += 8) {
    uint2 v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      const uint64_t idx = __umul64hi(s, words);  // uniform in [0, words)
      v[j] = __ldg(buf + idx);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) acc ^= v[j].x + v[j].y;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---------------------------------------------------------------------------
// Ray streams: n independent intersectOctree calls (coneTrace = false).
// ---------------------------------------------------------------------------
struct RayRec { float ox, oy, oz, dx, dy, dz; };
struct HitRec { uint32_t id; float t; uint32_t value, iter; };

template <bool FAST, bool WIDE = false>  // WIDE: 16-byte stack entries (SVO_OPT_STREAM_KERNEL 2)
__global__ void __launch_bounds__(128) k_cast_stream(SceneView sc, const RayRec *__restrict__ rays, const uint32_t *__restrict__ order,
                                                     uint64_t n, HitRec *__restrict__ out, int maxDepth) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = order ? (uint64_t)order[i] : i;
    const RayRec ray = rays[r];
    CastRes res;
    res.value = res.pointer = res.iter = res.depth = 0u;
    res.t = 0.0f; res.scale = 0.0f; res.dbg = 0.0f; res.dbg_init = 0;
    res.normal = mk3(0.f, 0.f, 0.f); res.voxelPos = mk3(0.f, 0.f, 0.f);
    uint32_t loops = 0;
    bool hit;
    if (WIDE) {
      uint4 wide[kMaxScale + 1];
      WideStack ws;
      ws.p = wide;
      hit = cast_ray_on<FAST, false, false, false>(ws, sc, mk3(ray.ox, ray.oy, ray.oz), mk3(ray.dx, ray.dy, ray.dz), maxDepth, false, 11, res, loops, nullptr, true);
    } else {
      hit = cast_ray<FAST>(sc, mk3(ray.ox, ray.oy, ray.oz), mk3(ray.dx, ray.dy, ray.dz), maxDepth, false, 11, res, loops);
    }
    HitRec h;
    h.id = hit ? res.pointer : kNoHit;
    h.t = hit ? res.t : 0.0f;
    h.value = hit ? res.value : 0u;
    h.iter = loops;
    out[r] = h;
  }
}

// ---------------------------------------------------------------------------
// Ray streams, persistent threads with warp-level ray fetch (Aila & Laine's while-while organisation; SVO_OPT_STREAM_KERNEL 1).
// Free-standing rays carry no shading state, so a lane that has finished can be re-armed for the price of
// Trav::setup: warps take chunks of the (binned) stream from a global counter, leave the traversal loop when
// kRefillIdle lanes are free, write those lanes' hit records and hand them the next rays.  Trip counts of incoherent
// rays range from 1 to 1500 within a warp; the grid-stride kernel above waits for the slowest ray of every 32.
// ---------------------------------------------------------------------------
constexpr unsigned kStreamChunk = 256;  // rays a warp takes from the stream per atomic

template <bool FAST>
__global__ void __launch_bounds__(128, 8) k_cast_stream_persistent(SceneView sc, const RayRec *__restrict__ rays, const uint32_t *__restrict__ order,
                                                                  uint64_t n, HitRec *__restrict__ out, int maxDepth,
                                                                  unsigned int *__restrict__ next_chunk) {
  const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
  uint4 wide[kMaxScale + 1];
  WideStack stk;
  stk.p = wide;
  Trav<FAST> T;
  uint64_t mine = 0, pool_next = 0, pool_end = 0;
  int busy = 0;
  bool dry = false;  // the stream has no chunks left
  for (;;) {
    unsigned idle = __ballot_sync(0xffffffffu, !busy);
    while (idle != 0u) {
      if (pool_next == pool_end) {
        if (dry) break;
        unsigned c = 0;
        if (lane == 0) c = atomicAdd(next_chunk, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        pool_next = (uint64_t)c * kStreamChunk;
        if (pool_next >= n) { dry = true; pool_next = pool_end = 0; break; }
        pool_end = pool_next + kStreamChunk < n ? pool_next + kStreamChunk : n;
      }
      const uint64_t avail = pool_end - pool_next;
      const unsigned rank = __popc(idle & lt_mask);
      if (!busy && rank < avail) {
        const uint64_t i = pool_next + rank;
        mine = order ? (uint64_t)order[i] : i;
        const RayRec ray = rays[mine];
        T.setup(sc, mk3(ray.ox, ray.oy, ray.oz), mk3(ray.dx, ray.dy, ray.dz), maxDepth, false, 11, nullptr);
        if (T.nan_ray(nullptr)) {  // ends before the loop: the record is final
          CastRes res;
          cast_res_clear(res);
          uint32_t loops = 0;
          T.finish(sc, TRAV_MISS, res, loops);
          HitRec h;
          h.id = kNoHit; h.t = 0.0f; h.value = 0u; h.iter = loops;
          out[mine] = h;
        } else {
          busy = 1;
        }
      }
      const unsigned take = (unsigned)__popc(idle);
      pool_next += avail < take ? avail : take;
      idle = __ballot_sync(0xffffffffu, !busy);
    }
    const int busy0 = __popc(__ballot_sync(0xffffffffu, busy));
    if (busy0 == 0) break;  // stream dry, every ray finished
    const int limit = dry && pool_next == pool_end ? 0 : max(busy0 - kRefillIdle, 0);
    for (;;) {
      if (busy) {
        int status = TRAV_CONTINUE;
#pragma unroll 1
        for (int k = 0; k < kRefillBatch && status == TRAV_CONTINUE; k++) status = T.step(sc, stk, nullptr);
        if (status != TRAV_CONTINUE) {
          busy = 0;
          CastRes res;
          cast_res_clear(res);
          uint32_t loops = 0;
          const bool hit = T.finish(sc, status, res, loops);
          HitRec h;
          h.id = hit ? res.pointer : kNoHit;
          h.t = hit ? res.t : 0.0f;
          h.value = hit ? res.value : 0u;
          h.iter = loops;
          out[mine] = h;
        }
      }
      if (__popc(__ballot_sync(0xffffffffu, busy)) <= limit) break;
    }
  }
}

// ---------------------------------------------------------------------------
// Beam pre-pass (reference src/shaders/svobeam.comp:617-636): one
// UN-normalised ray through pixel (4gx, 4gy); stores res.t (0 on miss, where
// upstream leaves it undefined).
// ---------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(128) k_beam(SceneView sc, FrameParams f, float *__restrict__ beam, int W, int H) {
  const int bw = W >> 2, bh = H >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const int gy = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
  if (gx >= bw || gy >= bh) return;
  const float fx = fdiv(fadd((float)(gx * 4), 0.5f), (float)W);
  const float fy = fdiv(fadd((float)(gy * 4), 0.5f), (float)H);
  vec3 dir;
  dir.x = mixf(mixf(f.l1[0], f.l2[0], fy), mixf(f.r1[0], f.r2[0], fy), fx);
  dir.y = mixf(mixf(f.l1[1], f.l2[1], fy), mixf(f.r1[1], f.r2[1], fy), fx);
  dir.z = mixf(mixf(f.l1[2], f.l2[2], fy), mixf(f.r1[2], f.r2[2], fy), fx);
  CastRes res;
  res.value = res.pointer = res.iter = res.depth = 0u;
  res.t = 0.0f; res.scale = 0.0f; res.dbg = 0.0f; res.dbg_init = 0;
  res.normal = mk3(0.f, 0.f, 0.f); res.voxelPos = mk3(0.f, 0.f, 0.f);
  uint32_t loops = 0;
  const bool hit = cast_ray<FAST>(sc, mk3(f.camPos[0], f.camPos[1], f.camPos[2]), dir, f.maxDepth, false, f.coneDepth, res, loops);
  beam[(size_t)gy * (size_t)bw + (size_t)gx] = hit ? res.t : 0.0f;
}

// ---------------------------------------------------------------------------
// Conservative beam pre-pass (what svobeam.comp:617-636 set out to do, done so that the frame cannot
// change).  Upstream casts ONE un-normalised ray through the first pixel of every 4x4 block and starts the block's 16
// fine rays at its hit distance: not a lower bound (a block's other pixels can hit earlier), in different units than
// the fine pass's normalised rays, undefined on a miss.  Here:
//  1. k_beam_lattice casts a NORMALISED ray through every corner of the 4x4 pixel grid ((W/4+1) x (H/4+1) lattice rays)
//     with an LOD stop: a non-empty node no larger than kappa x the lattice spacing at that distance counts as solid.
//     It reports t_entry - sqrt(3) * node size: no ray from the same origin can enter that node earlier.
//  2. k_beam_minfilter gives every block the minimum over the lattice rays within R lattice steps.  With kappa = 2*sqrt(2)
//     every reported node is at least sqrt(2) x the lattice spacing wide, so any non-empty node that meets a block's
//     frustum contains a lattice ray within 3.2 spacings of the block (convexity: its projection contains a disc of
//     diameter >= sqrt(2) d on the way from the block to its inscribed disc); R = 4 covers that.  So the block's value is
//     a true lower bound on the hit distance of each of its 16 pixels (+inf: they all miss), less a safety margin.
//  3. The fine pass (svo_frame.flags bit 1) starts the PRIMARY cast's walk at that distance (Trav::setup t_floor): the
//     cells skipped are empty, the first non-empty cell met is the one the full walk meets, entered from the same
//     empty cell: colour, depth and hit are unchanged bit for bit; only the iteration count falls.
// ---------------------------------------------------------------------------
constexpr float kBeamKappa = 2.8284271f;
constexpr int kBeamRadius = 4;

__device__ __forceinline__ vec3 beam_lattice_dir(const FrameParams &f, int gx, int gy, int W, int H) {
  const float fx = (float)(4 * gx) / (float)W, fy = (float)(4 * gy) / (float)H;
  vec3 d;
  d.x = mixf(mixf(f.l1[0], f.l2[0], fy), mixf(f.r1[0], f.r2[0], fy), fx);
  d.y = mixf(mixf(f.l1[1], f.l2[1], fy), mixf(f.r1[1], f.r2[1], fy), fx);
  d.z = mixf(mixf(f.l1[2], f.l2[2], fy), mixf(f.r1[2], f.r2[2], fy), fx);
  return normalize3(d);
}

// Lattice rows [row0, row1) only, every value stored into the dst.n lattice buffers of `dst` (this GPU's and, over NVLink, its
// peers': in the tile partition every rank traces 1/N of the lattice for everybody); the last CTA to leave bumps the fences
// of `sig` ("this rank's rows of the frame's lattice are stored").  `ticket`: two words, zero between launches.
__global__ void __launch_bounds__(128) k_beam_lattice(SceneView sc, FrameParams f, float *__restrict__ lattice, int W, int H, int row0, int row1,
                                                      FenceList dst, FenceList sig, unsigned int *__restrict__ ticket) {
  const int lw = (W >> 2) + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const int gy = row0 + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
  const bool live = gx < lw && gy < row1;
  float out_value = 0.0f;
  if (live) {
  const vec3 d0 = beam_lattice_dir(f, gx, gy, W, H), dx = beam_lattice_dir(f, gx + 1, gy, W, H), dy = beam_lattice_dir(f, gx, gy + 1, W, H);
  const vec3 ex = mk3(dx.x - d0.x, dx.y - d0.y, dx.z - d0.z), ey = mk3(dy.x - d0.x, dy.y - d0.y, dy.z - d0.z);
  // lattice spacing per unit distance around this ray, with slack for its variation over the filter window
  const float spacing = 1.25f * sqrtf(fmaxf(ex.x * ex.x + ex.y * ex.y + ex.z * ex.z, ey.x * ey.x + ey.y * ey.y + ey.z * ey.z));
  const float inf = __uint_as_float(0x7f800000u);
  float out = inf;
  Trav<true> T;
  uint2 stk[kMaxScale + 1];
  T.setup(sc, mk3(f.camPos[0], f.camPos[1], f.camPos[2]), d0, f.maxDepth, false, f.coneDepth, nullptr);
  if (!(d0.x == d0.x && d0.y == d0.y && d0.z == d0.z)) {
    out = 0.0f;  // degenerate camera: no bound
  } else if (!T.nan_ray(nullptr)) {
    for (;;) {
      const uint32_t m = T.pd.y >> (T.idx ^ T.oct);
      if ((m & 0x10000u) != 0u && T.t_min <= T.t_max && T.scale_exp2 <= kBeamKappa * spacing * T.t_min) {  // LOD stop
        out = T.t_min - 1.7320508f * T.scale_exp2;
        break;
      }
      const int st = T.step(sc, stk, nullptr);
      if (st == TRAV_HIT) { out = T.t_min - 1.7320508f * T.scale_exp2; break; }
      if (st == TRAV_MISS) {
        if (T.iter > (float)kMaxIterations) out = 0.0f;  // the walk was cut short by the iteration cap: no bound
        break;
      }
    }
  } else {
    out = 0.0f;
  }
  out_value = fmaxf(out, 0.0f);
  }
  if (live) {
    const size_t at = (size_t)gy * (size_t)lw + (size_t)gx;
    if (dst.n == 0) lattice[at] = out_value;
    for (int i = 0; i < dst.n; i++) ((float *)dst.p[i])[at] = out_value;
  }
  if (sig.n > 0) {  // last CTA out: everything this launch stored is visible system-wide before the bumps
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned t = atomicAdd(ticket, 1u);
      if (t == gridDim.x * gridDim.y - 1u) {
        ticket[0] = 0u;
        __threadfence_system();
        for (int i = 0; i < sig.n; i++) atomicAdd_system(sig.p[i], 1u);
      }
    }
  }
}

__global__ void __launch_bounds__(128) k_beam_minfilter(const float *__restrict__ lattice, float *__restrict__ beam, int W, int H) {
  const int bw = W >> 2, bh = H >> 2, lw = bw + 1, lh = bh + 1;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
  if (bx >= bw || by >= bh) return;
  float v = __uint_as_float(0x7f800000u);
  const int x0 = max(bx - kBeamRadius, 0), x1 = min(bx + 1 + kBeamRadius, lw - 1);
  const int y0 = max(by - kBeamRadius, 0), y1 = min(by + 1 + kBeamRadius, lh - 1);
  for (int y = y0; y <= y1; y++)
    for (int x = x0; x <= x1; x++) v = fminf(v, __ldg(lattice + (size_t)y * (size_t)lw + (size_t)x));
  // safety margin: the lattice pass contracts its t arithmetic into FMAs, the fine pass rounds every operation
  if (v < __uint_as_float(0x7f800000u)) v = fmaxf(v * 0.9999f - 1e-5f, 0.0f);
  beam[(size_t)by * (size_t)bw + (size_t)bx] = v;
}

__global__ void k_math_probe(int fn, const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r;
  switch (fn) {
    case 0: r = det_sin(x[i]); break;
    case 1: r = det_cos(x[i]); break;
    case 2: r = det_acos(x[i]); break;
    case 3: r = det_exp(x[i]); break;
    default: r = det_rand(x[i], y[i]); break;
  }
  out[i] = r;
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
// Variant 9 keeps kSmemStackLevels stack entries per thread in shared memory (scales 9..22).  A cast stops descending at
// maxDepth, or at coneDepth once the sticky cone cut has fired -- and if it is already deeper than coneDepth at that moment it
// runs down to the tree's leaves -- so all three must fit; otherwise the default kernel renders the frame.
static bool smem_stack_fits(const LaunchCfg &cfg, const FrameParams &f) {
  return f.maxDepth <= kSmemStackLevels && f.coneDepth <= kSmemStackLevels && cfg.scene_levels <= kSmemStackLevels;
}
static bool split_applies(const LaunchCfg &cfg, const FrameParams &f) {
  return (cfg.kernel == 15 || cfg.kernel == 16) && !cfg.fast && !cfg.aux && cfg.band_stride == 0 && f.renderMode == 0 && f.casts >= 2 && cfg.split.q[0] != nullptr;
}
int render_launches(const LaunchCfg &cfg, const FrameParams &f) { return split_applies(cfg, f) ? 2 : 1; }

cudaError_t launch_render(const LaunchCfg &cfg_in, const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H,
                          int y0, int y1, cudaStream_t stream) {
  LaunchCfg cfg = cfg_in;
  if (split_applies(cfg, f)) {
    if (y1 <= y0) return cudaSuccess;
    if ((uint64_t)W * (uint64_t)(y1 - y0) > cfg.split.capacity) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(cfg.split.counters, 0, 2 * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const dim3 grid((W + 15) / 16, (y1 - y0 + 7) / 8);
    const int pgrid = cfg.sm_count * 8;
#define SVO_LAUNCH_SPLIT(B, P)                                                                              \
  do {                                                                                                      \
    SVO_LAUNCH(grid, 128, stream, k_split_primary<B, P>)(sc, f, pl, W, H, y0, y1, cfg.split);                \
    SVO_LAUNCH(pgrid, 128, stream, k_split_bounce<B, P>)(sc, f, pl, W, cfg.split);                           \
  } while (0)
    if (cfg.kernel == 16) { if (cfg.box) SVO_LAUNCH_SPLIT(true, true); else SVO_LAUNCH_SPLIT(false, true); }
    else { if (cfg.box) SVO_LAUNCH_SPLIT(true, false); else SVO_LAUNCH_SPLIT(false, false); }
#undef SVO_LAUNCH_SPLIT
    return cudaGetLastError();
  }
  if (cfg.kernel == 15 || cfg.kernel == 16) cfg.kernel = 10;  // frames the split does not cover: the default kernel
  if (cfg.kernel == 1) {
    if (y1 <= y0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(cfg.tile_counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const int grid = cfg.sm_count * cfg.ctas_per_sm;
    if (cfg.fast) {
      if (cfg.aux) SVO_LAUNCH(grid, 128, stream, k_render_persistent<true, true>)(sc, f, pl, W, H, y0, y1, cfg.tile_counter);
      else SVO_LAUNCH(grid, 128, stream, k_render_persistent<true, false>)(sc, f, pl, W, H, y0, y1, cfg.tile_counter);
    } else {
      if (cfg.aux) SVO_LAUNCH(grid, 128, stream, k_render_persistent<false, true>)(sc, f, pl, W, H, y0, y1, cfg.tile_counter);
      else SVO_LAUNCH(grid, 128, stream, k_render_persistent<false, false>)(sc, f, pl, W, H, y0, y1, cfg.tile_counter);
    }
    return cudaGetLastError();
  }
  if (cfg.kernel == 6 && cfg.band_stride == 0) {
    const dim3 grid((W + 15) / 16, (y1 - y0 + 7) / 8);
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
#define SVO_LAUNCH_BINNED(F, A, B) SVO_LAUNCH(grid, 128, stream, k_render_tile_binned<F, A, B>)(sc, f, pl, W, H, y0, y1)
    if (cfg.fast) {
      if (cfg.aux) SVO_LAUNCH_BINNED(true, true, false);
      else if (cfg.box) SVO_LAUNCH_BINNED(true, false, true);
      else SVO_LAUNCH_BINNED(true, false, false);
    } else {
      if (cfg.aux) SVO_LAUNCH_BINNED(false, true, false);
      else if (cfg.box) SVO_LAUNCH_BINNED(false, false, true);
      else SVO_LAUNCH_BINNED(false, false, false);
    }
#undef SVO_LAUNCH_BINNED
    return cudaGetLastError();
  }
  if (cfg.kernel >= 18 && cfg.kernel <= 26 && !cfg.fast && !cfg.aux && cfg.band_stride == 0) {  // ablations of variant 13's parts (production instance only)
    const dim3 grid((W + 15) / 16, (y1 - y0 + 7) / 8);
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
#define SVO_LAUNCH_BALMASK(B, M) SVO_LAUNCH(grid, 128, stream, k_render_tile_balanced<false, B, M>)(sc, f, pl, W, H, y0, y1)
#define SVO_LAUNCH_BAL(B)                                  \
  do {                                                     \
    if (cfg.kernel == 18) SVO_LAUNCH_BALMASK(B, 13);       \
    else if (cfg.kernel == 19) SVO_LAUNCH_BALMASK(B, 7);   \
    else if (cfg.kernel == 20) SVO_LAUNCH_BALMASK(B, 1);   \
    else if (cfg.kernel == 21) SVO_LAUNCH_BALMASK(B, 5);   \
    else if (cfg.kernel == 22) SVO_LAUNCH(grid, 128, stream, k_render_tile_balanced<false, B, 15, 7>)(sc, f, pl, W, H, y0, y1); \
    else if (cfg.kernel == 23) SVO_LAUNCH(grid, 128, stream, k_render_tile_balanced<false, B, 1, 7>)(sc, f, pl, W, H, y0, y1);  \
    else if (cfg.kernel == 24) SVO_LAUNCH(grid, 128, stream, k_render_tile_balanced<false, B, 15, 6>)(sc, f, pl, W, H, y0, y1); \
    else if (cfg.kernel == 25) SVO_LAUNCH(grid, 128, stream, k_render_tile_balanced<false, B, 15, 9>)(sc, f, pl, W, H, y0, y1); \
    else SVO_LAUNCH(grid, 128, stream, k_render_tile_balanced<false, B, 15, 10>)(sc, f, pl, W, H, y0, y1);                      \
  } while (0)
    if (cfg.box) SVO_LAUNCH_BAL(true); else SVO_LAUNCH_BAL(false);
#undef SVO_LAUNCH_BAL
#undef SVO_LAUNCH_BALMASK
    return cudaGetLastError();
  }
  if (cfg.kernel >= 18 && cfg.kernel <= 26) cfg.kernel = 13;
  if (cfg.kernel == 9 && !smem_stack_fits(cfg, f)) cfg.kernel = 13;
  if (cfg.kernel >= 9 && cfg.kernel <= 13 && !cfg.fast && cfg.band_stride == 0 && (cfg.kernel != 9 || smem_stack_fits(cfg, f))) {
    const dim3 grid((W + 15) / 16, (y1 - y0 + 7) / 8);
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
#define SVO_LAUNCH_X(A, B)                                                                                           \
  do {                                                                                                               \
    if (cfg.kernel == 9) SVO_LAUNCH(grid, 128, stream, k_render_tile_stack<A, B, 2>)(sc, f, pl, W, H, y0, y1);        \
    else if (cfg.kernel == 10) SVO_LAUNCH(grid, 128, stream, k_render_tile_stack<A, B, 1>)(sc, f, pl, W, H, y0, y1);  \
    else if (cfg.kernel == 11) SVO_LAUNCH(grid, 128, stream, k_render_tile_regs<A, B, 7>)(sc, f, pl, W, H, y0, y1);   \
    else if (cfg.kernel == 12) SVO_LAUNCH(grid, 128, stream, k_render_tile_regs<A, B, 6>)(sc, f, pl, W, H, y0, y1);   \
    else SVO_LAUNCH(grid, 128, stream, k_render_tile_balanced<A, B>)(sc, f, pl, W, H, y0, y1);                        \
  } while (0)
    if (cfg.aux) SVO_LAUNCH_X(true, false);
    else if (cfg.box) SVO_LAUNCH_X(false, true);
    else SVO_LAUNCH_X(false, false);
#undef SVO_LAUNCH_X
    return cudaGetLastError();
  }
  if (cfg.kernel == 17 && !cfg.fast) {
    const int stride = cfg.band_stride > 0 ? cfg.band_stride : 1, offset = cfg.band_stride > 0 ? cfg.band_offset : 0;
    const int band_ctas = cfg.band_stride > 0 && cfg.band_ctas > 0 ? cfg.band_ctas : 1;
    const int bands = ((y1 - y0 + 7) / 8 + band_ctas - 1) / band_ctas;
    const unsigned nblocks_y = bands > offset ? (unsigned)(((bands - offset + stride - 1) / stride) * band_ctas) : 0u;
    const unsigned blocks_x = (unsigned)(W + 15) / 16u;
    const uint64_t nctas = (uint64_t)blocks_x * nblocks_y;  // one 16x8 block = four tiles = one CTA's worth of warps
    const uint64_t cap = (uint64_t)cfg.sm_count * 8u;
    const unsigned grid = (unsigned)(nctas < cap ? nctas : cap);
    if (grid == 0) {  // nothing to render: the fences of the frame are still owed
      if (cfg.fences.n > 0) SVO_LAUNCH(1, 32, stream, k_fence_signal)(cfg.fences);
      return cudaGetLastError();
    }
    if (cfg.aux) SVO_LAUNCH(grid, 128, stream, k_render_tile_queue<true, false>)(sc, f, pl, W, H, y0, y1, stride, offset, band_ctas, nblocks_y, cfg.tile_queue, cfg.fences);
    else if (cfg.box) SVO_LAUNCH(grid, 128, stream, k_render_tile_queue<false, true>)(sc, f, pl, W, H, y0, y1, stride, offset, band_ctas, nblocks_y, cfg.tile_queue, cfg.fences);
    else SVO_LAUNCH(grid, 128, stream, k_render_tile_queue<false, false>)(sc, f, pl, W, H, y0, y1, stride, offset, band_ctas, nblocks_y, cfg.tile_queue, cfg.fences);
    return cudaGetLastError();
  }
  if ((cfg.kernel == 7 || cfg.kernel == 8) && cfg.band_stride == 0) {
    // a CTA covers 32 x 16 pixels (variant 7: 4 pixels per thread) or 32 x 8 (variant 8: 2 per thread)
    const dim3 grid((W + 31) / 32, cfg.kernel == 7 ? (y1 - y0 + 15) / 16 : (y1 - y0 + 7) / 8);
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
#define SVO_LAUNCH_REFILL(F, A, B)                                                                                    \
  do {                                                                                                                \
    if (cfg.kernel == 7) SVO_LAUNCH(grid, 128, stream, k_render_tile_refill<F, A, B, 4>)(sc, f, pl, W, H, y0, y1);     \
    else SVO_LAUNCH(grid, 128, stream, k_render_tile_refill<F, A, B, 2>)(sc, f, pl, W, H, y0, y1);                     \
  } while (0)
    if (cfg.fast) {
      if (cfg.aux) SVO_LAUNCH_REFILL(true, true, false);
      else if (cfg.box) SVO_LAUNCH_REFILL(true, false, true);
      else SVO_LAUNCH_REFILL(true, false, false);
    } else {
      if (cfg.aux) SVO_LAUNCH_REFILL(false, true, false);
      else if (cfg.box) SVO_LAUNCH_REFILL(false, false, true);
      else SVO_LAUNCH_REFILL(false, false, false);
    }
#undef SVO_LAUNCH_REFILL
    return cudaGetLastError();
  }
  if (cfg.kernel == 4 && !cfg.aux && !cfg.fast && cfg.box && cfg.band_stride == 0) {
    const dim3 grid((W + 15) / 16, (y1 - y0 + 7) / 8);
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
    SVO_LAUNCH(grid, 128, stream, k_render_tile_smem)(sc, f, pl, W, H, y0, y1);
    return cudaGetLastError();
  }
  const dim3 block(cfg.kernel == 5 ? 64 : 128);
  const int stride = cfg.band_stride > 0 ? cfg.band_stride : 1, offset = cfg.band_stride > 0 ? cfg.band_offset : 0;
  const int band_ctas = cfg.band_stride > 0 && cfg.band_ctas > 0 ? cfg.band_ctas : 1;
  const int bands = ((y1 - y0 + 7) / 8 + band_ctas - 1) / band_ctas;  // bands of band_ctas CTA rows (the last may be short)
  const bool small_cta = cfg.kernel == 5;  // experiment: 64-thread CTAs (8x8 pixels)
  const dim3 grid(small_cta ? (W + 7) / 8 : (W + 15) / 16, bands > offset ? ((bands - offset + stride - 1) / stride) * band_ctas : 0);
  if (grid.x == 0 || grid.y == 0) return cudaSuccess;
#define SVO_LAUNCH_TILE(F, A, B) SVO_LAUNCH(grid, block, stream, k_render_tile<F, A, B>)(sc, f, pl, W, H, y0, y1, stride, offset, band_ctas)
#define SVO_LAUNCH_TILE_WIDE(A, B) SVO_LAUNCH(grid, block, stream, k_render_tile<false, A, B, 1>)(sc, f, pl, W, H, y0, y1, stride, offset, band_ctas)
  if (cfg.kernel == 14 && !cfg.fast) {  // variant 10's stack entries in the band-interleaved launch (multi-GPU tile partition); not yet measured
    if (cfg.aux) SVO_LAUNCH_TILE_WIDE(true, false);
    else if (cfg.box) SVO_LAUNCH_TILE_WIDE(false, true);
    else SVO_LAUNCH_TILE_WIDE(false, false);
  } else if (cfg.fast) {
    if (cfg.aux) SVO_LAUNCH_TILE(true, true, false);
    else if (cfg.box) SVO_LAUNCH_TILE(true, false, true);
    else SVO_LAUNCH_TILE(true, false, false);
  } else {
    if (cfg.aux) SVO_LAUNCH_TILE(false, true, false);
    else if (cfg.box) SVO_LAUNCH_TILE(false, false, true);
    else SVO_LAUNCH_TILE(false, false, false);
  }
#undef SVO_LAUNCH_TILE
#undef SVO_LAUNCH_TILE_WIDE
  return cudaGetLastError();
}

cudaError_t launch_render_stats(const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H, int y0, int y1,
                                unsigned long long *d_counters, cudaStream_t stream, bool executed) {
  const dim3 block(128);
  const dim3 grid((W + 15) / 16, (y1 - y0 + 7) / 8);
  if (grid.x == 0 || grid.y == 0) return cudaSuccess;
  if (executed) SVO_LAUNCH(grid, block, stream, k_render_stats<true>)(sc, f, pl, W, H, y0, y1, d_counters);
  else SVO_LAUNCH(grid, block, stream, k_render_stats<false>)(sc, f, pl, W, H, y0, y1, d_counters);
  return cudaGetLastError();
}

cudaError_t launch_fence_signal(const FenceList &fl, cudaStream_t stream) {
  SVO_LAUNCH(1, 32, stream, k_fence_signal)(fl);
  return cudaGetLastError();
}
cudaError_t launch_fence_wait(unsigned int *fence, unsigned int *dead, unsigned int target, cudaStream_t stream) {
  SVO_LAUNCH(1, 1, stream, k_fence_wait)(fence, dead, target);
  return cudaGetLastError();
}

cudaError_t launch_fence_wait_signal(unsigned int *fence, unsigned int *dead, unsigned int target, const FenceList &fl, cudaStream_t stream) {
  SVO_LAUNCH(1, 32, stream, k_fence_wait_signal)(fence, dead, target, fl);
  return cudaGetLastError();
}

cudaError_t launch_gather_probe(const void *buf, uint64_t words, int loads, int blocks, uint32_t *sink, cudaStream_t stream) {
  SVO_LAUNCH(blocks, 256, stream, k_gather_probe)((const uint2 *)buf, words, loads, sink);
  return cudaGetLastError();
}

cudaError_t launch_cast(const LaunchCfg &cfg, const SceneView &sc, const void *d_rays, const uint32_t *d_order, uint64_t n,
                        void *d_out, int maxDepth, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (cfg.stream_kernel == 1) {  // persistent threads with warp-level ray fetch
    cudaError_t e = cudaMemsetAsync(cfg.tile_counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const uint64_t warps = (n + 31) / 32;
    const uint64_t ctas = (warps + 3) / 4, cap_ctas = (uint64_t)cfg.sm_count * 8u;
    const unsigned pgrid = (unsigned)(ctas < cap_ctas ? ctas : cap_ctas);
    if (cfg.fast) SVO_LAUNCH(pgrid, 128, stream, k_cast_stream_persistent<true>)(sc, (const RayRec *)d_rays, d_order, n, (HitRec *)d_out, maxDepth, cfg.tile_counter);
    else SVO_LAUNCH(pgrid, 128, stream, k_cast_stream_persistent<false>)(sc, (const RayRec *)d_rays, d_order, n, (HitRec *)d_out, maxDepth, cfg.tile_counter);
    return cudaGetLastError();
  }
  const int block = 128;
  uint64_t want = (n + block - 1) / block;
  const uint64_t cap = (uint64_t)cfg.sm_count * 16u * 8u;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (cfg.stream_kernel == 2 && !cfg.fast) SVO_LAUNCH(grid, block, stream, k_cast_stream<false, true>)(sc, (const RayRec *)d_rays, d_order, n, (HitRec *)d_out, maxDepth);
  else if (cfg.fast) SVO_LAUNCH(grid, block, stream, k_cast_stream<true>)(sc, (const RayRec *)d_rays, d_order, n, (HitRec *)d_out, maxDepth);
  else SVO_LAUNCH(grid, block, stream, k_cast_stream<false>)(sc, (const RayRec *)d_rays, d_order, n, (HitRec *)d_out, maxDepth);
  return cudaGetLastError();
}

cudaError_t launch_beam(const LaunchCfg &cfg, const SceneView &sc, const FrameParams &f, float *beam, int W, int H,
                        cudaStream_t stream) {
  const int bw = W >> 2, bh = H >> 2;
  if (bw == 0 || bh == 0) return cudaSuccess;
  const dim3 block(128), grid((bw + 15) / 16, (bh + 7) / 8);
  if (cfg.fast) SVO_LAUNCH(grid, block, stream, k_beam<true>)(sc, f, beam, W, H);
  else SVO_LAUNCH(grid, block, stream, k_beam<false>)(sc, f, beam, W, H);
  return cudaGetLastError();
}

cudaError_t launch_beam_lattice_rows(const SceneView &sc, const FrameParams &f, float *lattice, int W, int H, int row0, int row1,
                                     const FenceList &dst, const FenceList &sig, unsigned int *ticket, cudaStream_t stream) {
  const int bw = W >> 2, bh = H >> 2;
  if (bw == 0 || bh == 0) return cudaSuccess;
  if (row1 > bh + 1) row1 = bh + 1;
  if (row0 < 0) row0 = 0;
  if (row0 >= row1) {  // nothing to trace: the fences of the frame are still owed
    if (sig.n > 0) SVO_LAUNCH(1, 32, stream, k_fence_signal)(sig);
    return cudaGetLastError();
  }
  const dim3 lgrid((bw + 1 + 15) / 16, (row1 - row0 + 7) / 8);
  SVO_LAUNCH(lgrid, 128, stream, k_beam_lattice)(sc, f, lattice, W, H, row0, row1, dst, sig, ticket);
  return cudaGetLastError();
}
cudaError_t launch_beam_filter(const float *lattice, float *beam, int W, int H, cudaStream_t stream) {
  const int bw = W >> 2, bh = H >> 2;
  if (bw == 0 || bh == 0) return cudaSuccess;
  const dim3 fgrid((bw + 127) / 128, bh);
  SVO_LAUNCH(fgrid, 128, stream, k_beam_minfilter)(lattice, beam, W, H);
  return cudaGetLastError();
}
cudaError_t launch_beam_conservative(const SceneView &sc, const FrameParams &f, float *lattice, float *beam, int W, int H, cudaStream_t stream) {
  FenceList none;
  none.n = 0;
  for (int i = 0; i < 16; i++) none.p[i] = nullptr;
  cudaError_t e = launch_beam_lattice_rows(sc, f, lattice, W, H, 0, (H >> 2) + 1, none, none, nullptr, stream);
  if (e != cudaSuccess) return e;
  return launch_beam_filter(lattice, beam, W, H, stream);
}

cudaError_t launch_math_probe(int fn, const float *x, const float *y, float *out, uint64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  SVO_LAUNCH((unsigned)((n + 255) / 256), 256, stream, k_math_probe)(fn, x, y, out, n);
  return cudaGetLastError();
}

}  // namespace svo
