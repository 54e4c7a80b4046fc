// svo_wavefront.cu -- kernel variant 2: wavefront rendering (sm_100a).
//
// Same function as the tile kernel (svo_kernels.cu variant 0) -- reference
// src/shaders/svotrace.comp main/trace/intersectOctree -- organised so that the
// two very different kinds of work never share a warp:
//
//   k_wf_trav   persistent threads, warp-level ray fetch.  Only the traversal
//               loop (Trav::step).  A warp pulls 32 rays at a time from a global
//               queue (primary rays: one 8x4 pixel tile, generated in registers;
//               later casts: the compacted ray queue the shade kernel wrote) and
//               leaves its loop whenever kRefill lanes have finished, so those
//               lanes are re-armed instead of idling behind the slowest ray
//               (iteration counts run from 1 to 1500).  Re-arming costs a ray
//               load + Trav::setup, nothing else.
//   k_wf_shade  one thread per finished cast, all lanes busy: the code after
//               the loop (finish_hit), the shading between casts
//               (pixel_after_cast: RNG, ONB, sin/cos, materials, fog), the final
//               stores, and warp-aggregated compaction of the next cast's rays
//               (only pixels that still want a cast enter the next queue).
//
// Between the kernels a cast is 32 B of ray, 32 B of hit state and 96 B of
// pixel state in structure-of-array uint4 planes (coalesced 128-bit accesses).
#include "svo_kernels.h"
#include "svo_trace.cuh"

#ifndef SVO_LAUNCH  // see svo_kernels.cu
#ifdef SVO_HOST_EMU
#define SVO_LAUNCH(grid, block, stream, ...) simt::launcher(grid, block, __VA_ARGS__)
#else
#define SVO_LAUNCH(grid, block, stream, ...) __VA_ARGS__<<<grid, block, 0, stream>>>
#endif
#endif

namespace svo {

constexpr int kRefill = 8;  // a warp re-arms its idle lanes once this many have finished

SVO_DI void slot_to_pixel(unsigned slot, int tiles_x, int y0, int &x, int &y) {
  const unsigned tile = slot >> 5, k = slot & 31u;  // 8x4 pixel tiles, row-major tile order
  x = (int)(tile % (unsigned)tiles_x) * 8 + (int)(k & 7u);
  y = y0 + (int)(tile / (unsigned)tiles_x) * 4 + (int)(k >> 3);
}

template <bool FAST, bool PRIMARY>
__global__ void __launch_bounds__(128) k_wf_trav(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1,
                                                 const uint4 *__restrict__ rayA, const uint4 *__restrict__ rayB,
                                                 const unsigned *__restrict__ n_rays, uint4 *__restrict__ hitA,
                                                 uint4 *__restrict__ hitB, unsigned *__restrict__ work_counter) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int tiles_x = (W + 7) >> 3;
  const unsigned n = PRIMARY ? (unsigned)tiles_x * (unsigned)((y1 - y0 + 3) >> 2) * 32u : __ldg(n_rays);

  uint2 stk[kMaxScale + 1];
  Trav<FAST> T;
  bool busy = false;
  unsigned slot = 0;
  unsigned pool_next = 0, pool_end = 0;  // the warp's current batch of queue slots
  bool empty = false;

  for (;;) {
    // ---- warp-level ray fetch: re-arm idle lanes -----------------------------------
    unsigned want = __ballot_sync(0xffffffffu, !busy);
    while (want != 0u) {
      if (pool_next == pool_end) {
        if (empty) break;
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(work_counter, 32u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n) { empty = true; break; }
        pool_next = t;
        pool_end = min(t + 32u, n);
      }
      const unsigned avail = pool_end - pool_next;
      const unsigned rank = __popc(want & lt_mask);
      if (((want >> lane) & 1u) != 0u && rank < avail) {
        const unsigned s = pool_next + rank;
        if (PRIMARY) {
          int x, y;
          slot_to_pixel(s, tiles_x, y0, x, y);
          if (x < W && y < y1) {
            Pixel P;
            if (pixel_begin(f, pl, W, H, x, y, P)) {
              T.setup(sc, P.origin, P.dir, f.maxDepth, false, f.coneDepth, nullptr);
              busy = true;
              slot = s;
            }
          }
        } else {
          const uint4 a = __ldg(rayA + s), b = __ldg(rayB + s);
          T.setup(sc, mk3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)),
                  mk3(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y)), f.maxDepth, b.z != 0u, f.coneDepth,
                  nullptr);
          busy = true;
          slot = s;
        }
      }
      pool_next += min(avail, (unsigned)__popc(want));
      want = __ballot_sync(0xffffffffu, !busy);
    }
    const int busy0 = __popc(__ballot_sync(0xffffffffu, busy));
    if (busy0 == 0) break;  // queue drained and nothing in flight
    // ---- traverse until kRefill lanes are done (or, with the queue empty, all are) ----
    for (;;) {
      if (busy) {
        const int status = T.step(sc, stk, nullptr);
        if (status != TRAV_CONTINUE) {
          const HitState hs = T.export_hit(status);
          hitA[slot] = make_uint4(hs.pidx, hs.meta, hs.ipx, hs.ipy);
          hitB[slot] = make_uint4(hs.ipz, __float_as_uint(hs.t_min), hs.iter, 0u);
          busy = false;
        }
      }
      const int nb = __popc(__ballot_sync(0xffffffffu, busy));
      if (nb == 0 || (!empty && nb + kRefill <= busy0)) break;
    }
  }
}

struct WfQueues {
  const uint4 *hitA, *hitB;
  const uint4 *rayA_in, *rayB_in;
  const uint4 *st_in[6];
  const unsigned *n_in;
  uint4 *rayA_out, *rayB_out;
  uint4 *st_out[6];
  unsigned *n_out;
};

SVO_DI void pixel_unpack(const WfQueues &q, unsigned i, Pixel &P) {
  const uint4 ra = __ldg(q.rayA_in + i), rb = __ldg(q.rayB_in + i);
  const uint4 s0 = __ldg(q.st_in[0] + i), s1 = __ldg(q.st_in[1] + i), s2 = __ldg(q.st_in[2] + i);
  const uint4 s3 = __ldg(q.st_in[3] + i), s4 = __ldg(q.st_in[4] + i), s5 = __ldg(q.st_in[5] + i);
  P.origin = mk3(__uint_as_float(ra.x), __uint_as_float(ra.y), __uint_as_float(ra.z));
  P.dir = mk3(__uint_as_float(ra.w), __uint_as_float(rb.x), __uint_as_float(rb.y));
  P.cone = rb.z != 0u;
  P.x = (int)(s0.x & 0xFFFFu);
  P.y = (int)(s0.x >> 16);
  P.cast_i = (int)s0.y;
  P.depth = __uint_as_float(s0.z);
  P.beamDist = __uint_as_float(s0.w);
  P.acc = mk3(__uint_as_float(s1.x), __uint_as_float(s1.y), __uint_as_float(s1.z));
  P.mask = mk3(__uint_as_float(s1.w), __uint_as_float(s2.x), __uint_as_float(s2.y));
  P.res.value = s2.z;
  P.res.pointer = s2.w;
  P.res.iter = s3.x;
  P.res.depth = s3.y & 0x7FFFFFFFu;
  P.res.dbg_init = (int)(s3.y >> 31);
  P.res.t = __uint_as_float(s3.z);
  P.res.scale = __uint_as_float(s3.w);
  P.res.normal = mk3(__uint_as_float(s4.x), __uint_as_float(s4.y), __uint_as_float(s4.z));
  P.res.dbg = __uint_as_float(s4.w);
  P.res.voxelPos = mk3(__uint_as_float(s5.x), __uint_as_float(s5.y), __uint_as_float(s5.z));
  P.color = mk3(0.0f, 0.0f, 0.0f);
  P.hit_id = kNoHit;  // the primary-cast planes were stored by stage 0
  P.iter = 0;
  P.primary_t = 0.0f;
}

SVO_DI void pixel_pack(const WfQueues &q, unsigned j, const Pixel &P) {
  q.rayA_out[j] = make_uint4(__float_as_uint(P.origin.x), __float_as_uint(P.origin.y), __float_as_uint(P.origin.z), __float_as_uint(P.dir.x));
  q.rayB_out[j] = make_uint4(__float_as_uint(P.dir.y), __float_as_uint(P.dir.z), P.cone ? 1u : 0u, 0u);
  q.st_out[0][j] = make_uint4((uint32_t)P.x | ((uint32_t)P.y << 16), (uint32_t)P.cast_i, __float_as_uint(P.depth), __float_as_uint(P.beamDist));
  q.st_out[1][j] = make_uint4(__float_as_uint(P.acc.x), __float_as_uint(P.acc.y), __float_as_uint(P.acc.z), __float_as_uint(P.mask.x));
  q.st_out[2][j] = make_uint4(__float_as_uint(P.mask.y), __float_as_uint(P.mask.z), P.res.value, P.res.pointer);
  q.st_out[3][j] = make_uint4(P.res.iter, (P.res.depth & 0x7FFFFFFFu) | ((uint32_t)(P.res.dbg_init != 0) << 31), __float_as_uint(P.res.t),
                              __float_as_uint(P.res.scale));
  q.st_out[4][j] = make_uint4(__float_as_uint(P.res.normal.x), __float_as_uint(P.res.normal.y), __float_as_uint(P.res.normal.z),
                              __float_as_uint(P.res.dbg));
  q.st_out[5][j] = make_uint4(__float_as_uint(P.res.voxelPos.x), __float_as_uint(P.res.voxelPos.y), __float_as_uint(P.res.voxelPos.z), 0u);
}

// One thread per finished cast.  STAGE0: casts are the primary rays (slot = tile-ordered pixel).
template <bool AUX, bool STAGE0>
__global__ void __launch_bounds__(128) k_wf_shade(SceneView sc, FrameParams f, Planes pl, int W, int H, int y0, int y1, WfQueues q) {
  const unsigned lane = threadIdx.x & 31u;
  const int tiles_x = (W + 7) >> 3;
  const unsigned n = STAGE0 ? (unsigned)tiles_x * (unsigned)((y1 - y0 + 3) >> 2) * 32u : __ldg(q.n_in);
  const unsigned stride = gridDim.x * blockDim.x;
  // whole warps iterate together (the compaction below is warp-wide)
  for (unsigned base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += stride) {
    const unsigned i = base + lane;
    bool valid = i < n, more = false;
    Pixel P;
    if (STAGE0 && valid) {
      int x, y;
      slot_to_pixel(i, tiles_x, y0, x, y);
      valid = x < W && y < y1;
      if (valid) {
        if (!pixel_begin(f, pl, W, H, x, y, P)) {  // no cast at all (mode 4)
          pixel_store<AUX>(sc, f, pl, W, P);
          valid = false;
        }
      }
    } else if (valid) {
      pixel_unpack(q, i, P);
    }
    if (valid) {
      const uint4 ha = __ldg(q.hitA + i), hb = __ldg(q.hitB + i);
      HitState hs;
      hs.pidx = ha.x; hs.meta = ha.y; hs.ipx = ha.z; hs.ipy = ha.w;
      hs.ipz = hb.x; hs.t_min = __uint_as_float(hb.y); hs.iter = hb.z;
      more = pixel_finish_cast(sc, f, P, hs);
      if (STAGE0 && AUX) {  // the primary-cast planes
        const size_t p = (size_t)P.y * (size_t)W + (size_t)P.x;
        pl.hit_id[p] = P.hit_id;
        pl.iter[p] = P.iter;
        pl.primary_t[p] = P.primary_t;
      }
      if (!more) {  // the end of main() (:696-727)
        const size_t p = (size_t)P.y * (size_t)W + (size_t)P.x;
        pixel_finish(sc, f, pl, p, P);
        pl.rgba8[p] = make_uchar4(quant8(P.color.x), quant8(P.color.y), quant8(P.color.z), 255);
        pl.depth[p] = P.depth;
        if (AUX) pl.radiance[p] = make_float4(P.color.x, P.color.y, P.color.z, 1.0f);
      }
    }
    // warp-aggregated append of the pixels that cast again
    const unsigned m = __ballot_sync(0xffffffffu, more);
    if (m != 0u) {
      unsigned j0 = 0;
      if (lane == (unsigned)(__ffs(m) - 1)) j0 = atomicAdd(q.n_out, (unsigned)__popc(m));
      j0 = __shfl_sync(0xffffffffu, j0, __ffs(m) - 1);
      if (more) pixel_pack(q, j0 + __popc(m & ((1u << lane) - 1u)), P);
    }
  }
}

// ---------------------------------------------------------------------------
cudaError_t launch_render_wavefront(const LaunchCfg &cfg, const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H,
                                    int y0, int y1, const WaveWorkspace &ws, cudaStream_t stream) {
  if (y1 <= y0) return cudaSuccess;
  const int mode = f.renderMode;
  int stages = 0;  // number of casts a pixel can issue
  if (mode == 0) stages = f.casts;
  else if (mode == 2) stages = 2;
  else if (mode == 1 || mode == 3) stages = 1;
  if (stages > kWaveMaxStages) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(ws.counters, 0, sizeof(unsigned) * 2 * (kWaveMaxStages + 1), stream);
  if (e != cudaSuccess) return e;
  unsigned *work = ws.counters;                         // [stage] traversal work counters
  unsigned *count = ws.counters + kWaveMaxStages + 1;   // [stage] rays in the queue feeding stage s (s >= 1)
  const int trav_grid = cfg.sm_count * cfg.ctas_per_sm;
  const int shade_grid = cfg.sm_count * 8;
  for (int s = 0; s < (stages > 0 ? stages : 1); s++) {
    const int qi = (s + 1) & 1, qo = s & 1;  // stage s reads queue qi (written by stage s-1) and writes queue qo
    if (stages > 0) {
      if (s == 0) {
        if (cfg.fast) SVO_LAUNCH(trav_grid, 128, stream, k_wf_trav<true, true>)(sc, f, pl, W, H, y0, y1, nullptr, nullptr, nullptr, ws.hitA, ws.hitB, work + 0);
        else SVO_LAUNCH(trav_grid, 128, stream, k_wf_trav<false, true>)(sc, f, pl, W, H, y0, y1, nullptr, nullptr, nullptr, ws.hitA, ws.hitB, work + 0);
      } else {
        if (cfg.fast) SVO_LAUNCH(trav_grid, 128, stream, k_wf_trav<true, false>)(sc, f, pl, W, H, y0, y1, ws.rayA[qi], ws.rayB[qi], count + s, ws.hitA, ws.hitB, work + s);
        else SVO_LAUNCH(trav_grid, 128, stream, k_wf_trav<false, false>)(sc, f, pl, W, H, y0, y1, ws.rayA[qi], ws.rayB[qi], count + s, ws.hitA, ws.hitB, work + s);
      }
    }
    WfQueues q;
    q.hitA = ws.hitA; q.hitB = ws.hitB;
    q.rayA_in = ws.rayA[qi]; q.rayB_in = ws.rayB[qi];
    q.rayA_out = ws.rayA[qo]; q.rayB_out = ws.rayB[qo];
    for (int k = 0; k < 6; k++) { q.st_in[k] = ws.state[qi][k]; q.st_out[k] = ws.state[qo][k]; }
    q.n_in = count + s;
    q.n_out = count + s + 1;
    if (s == 0) {
      if (cfg.aux) SVO_LAUNCH(shade_grid, 128, stream, k_wf_shade<true, true>)(sc, f, pl, W, H, y0, y1, q);
      else SVO_LAUNCH(shade_grid, 128, stream, k_wf_shade<false, true>)(sc, f, pl, W, H, y0, y1, q);
    } else {
      if (cfg.aux) SVO_LAUNCH(shade_grid, 128, stream, k_wf_shade<true, false>)(sc, f, pl, W, H, y0, y1, q);
      else SVO_LAUNCH(shade_grid, 128, stream, k_wf_shade<false, false>)(sc, f, pl, W, H, y0, y1, q);
    }
  }
  return cudaGetLastError();
}

int wavefront_launches(const FrameParams &f) {
  const int mode = f.renderMode;
  int stages = mode == 0 ? f.casts : (mode == 2 ? 2 : ((mode == 1 || mode == 3) ? 1 : 0));
  return stages > 0 ? 2 * stages : 1;
}

}  // namespace svo
