// svo_kernels.h -- host-visible types and launchers of the trace kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace svo {

struct SceneView {
  const uint2 *desc;        // interior-node descriptors, BFS order, [ndesc]
  const uint32_t *refbase;  // reference byte offset of each node's child block, [ndesc]
  const uint8_t *raw;       // the reference node stream, [nbytes]
  uint64_t nbytes;
  uint32_t ndesc;
  uint32_t first_word_zero;  // octreeBuffer[0] == 0 (svotrace.comp:696)
  uint32_t zero, one, two, four, exp_unit;  // hold 0, 1, 2, 4, 1 << 23: operands ptxas cannot fold (imad() in svo_trace.cuh)
  const uint2 *top;  // shared-memory copy of desc[0, ntop) (kernel variant 4 only)
  uint32_t ntop;
  float box_lo[3], box_hi[3];  // padded bounds of everything a cast of this frame can hit (cube coordinates [1,2])
};

struct FrameParams {  // == svo_frame (include/svo_b200.h)
  float camPos[3];
  float l1[3], l2[3], r1[3], r2[3];
  int frameNumber, renderMode, useBeam, maxDepth, casts, coneDepth, mirrorValue, flags;
};

struct Planes {
  uchar4 *rgba8;
  float *depth;
  const float *beam;
  uint32_t *hit_id;
  uint32_t *iter;
  float *primary_t;
  float4 *radiance;
};

// Device workspace of kernel variant 15 (primary kernel + persistent bounce kernel): one record per deferred cast in five
// (variant 16: six) uint4 planes of `capacity` entries (image pixels), and two counters (records queued, next chunk to hand out).
struct SplitQueue {
  uint4 *q[6];
  unsigned int *counters;
  uint64_t capacity;
};

struct FenceList { unsigned int *p[16]; int n; };  // fence words (possibly in peer memory) to bump by one

struct LaunchCfg {
  bool fast;     // Ops<true>: fma-contracted t arithmetic
  bool aux;      // also write hit_id / iter / primary_t / radiance
  bool box;      // end casts that are outside the content box (SceneView::box_*)
  int kernel;    // variant selector (SVO_OPT_KERNEL)
  int stream_kernel;  // ray streams: 0 grid-stride kernel, 1 persistent threads with warp-level ray fetch (SVO_OPT_STREAM_KERNEL)
  int sm_count;
  int scene_levels;  // octree levels of the uploaded scene (variant 9 needs them to fit its shared-memory stack)
  int band_stride, band_offset;  // tile kernel: interleaved bands (stride 0 = all bands)
  int band_ctas;                 // CTA rows (8 image rows each) per band
  int ctas_per_sm;             // persistent grid = sm_count * ctas_per_sm
  unsigned int *tile_counter;  // device word: the persistent kernel's tile queue head
  unsigned int *tile_queue;    // variant 17: two device words {next tile, CTAs that have left}, zero between launches
  FenceList fences;            // variant 17: frame-complete fences the launch signals when its last CTA leaves (n = 0: none)
  SplitQueue split;            // variant 15 only (q[0] == nullptr: not allocated)
};

constexpr int kWaveMaxStages = 64;  // casts per pixel the wavefront path supports (svo_frame.casts <= 64)

// Device workspace of the wavefront path (variant 2): two ray queues, one hit-state buffer,
// two pixel-state queues (all SoA uint4 planes of `slots` entries) and the stage counters.
struct WaveWorkspace {
  uint4 *rayA[2], *rayB[2];
  uint4 *hitA, *hitB;
  uint4 *state[2][6];
  unsigned *counters;  // 2 * (kWaveMaxStages + 1) words
  uint64_t slots;
};

cudaError_t launch_render_wavefront(const LaunchCfg &cfg, const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H,
                                    int y0, int y1, const WaveWorkspace &ws, cudaStream_t stream);
int wavefront_launches(const FrameParams &f);
cudaError_t launch_render(const LaunchCfg &cfg, const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H,
                          int y0, int y1, cudaStream_t stream);
// kernels launch_render() enqueues for this frame (variant 15: two)
int render_launches(const LaunchCfg &cfg, const FrameParams &f);
cudaError_t launch_render_stats(const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H, int y0, int y1,
                                unsigned long long *d_counters, cudaStream_t stream, bool executed = false);
cudaError_t launch_fence_signal(const FenceList &fl, cudaStream_t stream);
cudaError_t launch_fence_wait(unsigned int *fence, unsigned int *dead, unsigned int target, cudaStream_t stream);
cudaError_t launch_fence_wait_signal(unsigned int *fence, unsigned int *dead, unsigned int target, const FenceList &fl, cudaStream_t stream);
cudaError_t launch_gather_probe(const void *buf, uint64_t words, int loads, int blocks, uint32_t *sink, cudaStream_t stream);
cudaError_t launch_cast(const LaunchCfg &cfg, const SceneView &sc, const void *d_rays, const uint32_t *d_order, uint64_t n,
                        void *d_out, int maxDepth, cudaStream_t stream);
cudaError_t launch_beam(const LaunchCfg &cfg, const SceneView &sc, const FrameParams &f, float *beam, int W, int H,
                        cudaStream_t stream);
// conservative beam pre-pass: lattice = scratch of (W/4+1) x (H/4+1) floats, beam = the (W/4) x (H/4) beam plane
cudaError_t launch_beam_lattice_rows(const SceneView &sc, const FrameParams &f, float *lattice, int W, int H, int row0, int row1,
                                     const FenceList &dst, const FenceList &sig, unsigned int *ticket, cudaStream_t stream);
cudaError_t launch_beam_filter(const float *lattice, float *beam, int W, int H, cudaStream_t stream);
cudaError_t launch_beam_conservative(const SceneView &sc, const FrameParams &f, float *lattice, float *beam, int W, int H, cudaStream_t stream);
struct CellBox;
cudaError_t gpu_transcode(const uint8_t *d_raw, uint64_t nbytes, uint2 *desc, uint32_t *refbase, uint2 *meta, uint64_t cap, uint64_t *ndesc,
                          uint32_t *nlevels, CellBox *leaf_box, CellBox *depth_box, bool *overflow, cudaStream_t stream);
// incremental transcode (svo_upload_range): store + diff a byte range, then re-walk only the subtrees the changed bytes touch
cudaError_t gpu_diff_apply(uint8_t *d_raw, const uint8_t *d_fresh, uint64_t start, uint64_t end, uint64_t old_nbytes, uint8_t *d_bitmap,
                           uint64_t span[2], void *arena, size_t arena_bytes, cudaStream_t stream);
cudaError_t gpu_patch(const uint8_t *d_raw, uint64_t nbytes, const uint8_t *d_bitmap, uint64_t start, uint64_t end, const uint64_t span[2],
                      uint2 *desc, uint32_t *refbase, uint2 *meta, uint8_t *flag, uint64_t cap, uint64_t *ndesc, CellBox *leaf_box,
                      CellBox *depth_box, bool *fallback, uint64_t stats[3], void *arena, size_t arena_bytes, cudaStream_t stream);
size_t ray_sort_temp_bytes(uint64_t n);
cudaError_t launch_ray_sort(const void *d_rays, uint64_t n, uint32_t *keys, uint32_t *keys_alt, uint32_t *idx, uint32_t *order_out,
                            void *temp, size_t temp_bytes, int mode, cudaStream_t stream);
cudaError_t launch_math_probe(int fn, const float *x, const float *y, float *out, uint64_t n, cudaStream_t stream);

}  // namespace svo
