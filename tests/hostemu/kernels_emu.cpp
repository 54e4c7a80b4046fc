// kernels_emu.cpp -- TEST INFRASTRUCTURE: svo_kernels.cu (the __global__ functions and their launchers) compiled by
// g++ and run on the coroutine SIMT emulator.  launch_render() below is the product's own launcher -- same variant
// dispatch, same grid arithmetic -- with <<<>>> routed to simt::run_grid.  Nothing here ships.
#include "cuda_host_shim.h"
#include "simt_emu.h"

#define SVO_HOST_EMU 1
#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), cudaSuccess)
#define cudaGetLastError() cudaSuccess
#include "../../svo_raytracer_b200/csrc/svo_kernels.cu"
#undef cudaMemsetAsync
#undef cudaGetLastError

#include <algorithm>
#include <string>

#include "emu_scene.h"

using namespace svo;

int emu_launch_wavefront(const LaunchCfg &cfg, const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H, int y0, int y1);

static int g_last_launches = 0;

extern "C" {

unsigned int g_fence_words[2] = {0u, 0u};  // variant 17 bumps both when its last CTA leaves
unsigned int emu_fence_word(int i) { return g_fence_words[i & 1]; }
int emu_last_render_launches() { return g_last_launches; }  // kernels the last emu_launch_render enqueued (variant 15: 2)

// svo_render_rows through the product's launch_render on the emulator.  kernel = SVO_OPT_KERNEL, ctas = CTAs of the
// persistent variant (sm_count * ctas_per_sm on the device).
int emu_launch_render(const emu_scene *s, const FrameParams *f, int W, int H, int y0, int y1, int kernel, int box, int aux, const float *beam,
                      uint8_t *rgba8, float *depth, uint32_t *hit_id, uint32_t *iter, float *primary_t, float *radiance, int band_stride,
                      int band_offset, int band_rows, int ctas, int nthreads) {
  const SceneView sc = emu_view_of(s, f);
  Planes pl;
  pl.rgba8 = (uchar4 *)rgba8;
  pl.depth = depth;
  pl.beam = beam;
  pl.hit_id = hit_id;
  pl.iter = iter;
  pl.primary_t = primary_t;
  pl.radiance = (float4 *)radiance;
  unsigned int tile_counter = 0;
  LaunchCfg cfg;
  cfg.fast = false;
  cfg.aux = aux != 0;
  cfg.box = box != 0 && !aux && f->renderMode != 1;  // box_allowed() of svo_capi.cu
  cfg.kernel = kernel;
  cfg.stream_kernel = 0;
  cfg.sm_count = ctas;
  cfg.scene_levels = 8;  // the emulator's worlds are at most 128^3
  cfg.ctas_per_sm = 1;
  cfg.band_stride = band_stride;
  cfg.band_offset = band_offset;
  cfg.band_ctas = band_rows / 8;
  cfg.tile_counter = &tile_counter;
  unsigned int tile_queue[2] = {0u, 0u};  // variant 17: {next tile, CTAs that have left}
  cfg.tile_queue = tile_queue;
  cfg.fences.n = 0;
  for (int i = 0; i < 16; i++) cfg.fences.p[i] = nullptr;
  if (kernel == 17) { cfg.fences.n = 2; cfg.fences.p[0] = &g_fence_words[0]; cfg.fences.p[1] = &g_fence_words[1]; }
  std::vector<uint4> split_planes;
  unsigned int split_counters[2] = {0u, 0u};
  cfg.split = SplitQueue();
  if (kernel == 15 || kernel == 16) {  // ensure_split() of svo_capi.cu
    const uint64_t cap = (uint64_t)W * (uint64_t)H;
    split_planes.resize((size_t)(6 * cap));
    for (int k = 0; k < 6; k++) cfg.split.q[k] = split_planes.data() + (size_t)k * cap;
    cfg.split.counters = split_counters;
    cfg.split.capacity = cap;
  }
  g_last_launches = render_launches(cfg, *f);
  simt::g_os_threads = (kernel == 1 || kernel == 2 || kernel == 17) ? 1 : nthreads;  // persistent kernels: the queue is consumed by whichever block runs
  if (kernel == 2) return emu_launch_wavefront(cfg, sc, *f, pl, W, H, y0, y1);  // wavefront_emu.cpp
  const int rc = (int)launch_render(cfg, sc, *f, pl, W, H, y0, y1, nullptr);
  if (kernel == 17 && (tile_queue[0] != 0u || tile_queue[1] != 0u)) return 700;  // the last CTA out must leave the queue ready for the next launch
  return rc;
}

// svo_beam_conservative through the product's launcher: beam = (W/4) x (H/4) floats
int emu_beam_conservative(const emu_scene *s, const FrameParams *f, int W, int H, float *beam, int nthreads) {
  const SceneView sc = emu_view_of(s, nullptr);
  std::vector<float> lattice((size_t)(W / 4 + 1) * (size_t)(H / 4 + 1) + 4);
  simt::g_os_threads = nthreads;
  return (int)launch_beam_conservative(sc, *f, lattice.data(), beam, W, H, nullptr);
}

// svo_beam_lattice_rows x parts + svo_beam_filter (the tile partition's shared pre-pass): part r traces lattice rows
// [r*lh/parts, (r+1)*lh/parts) into TWO lattice buffers (its own and a "peer's") and bumps two fence words from its last CTA.
// Returns 0 if both lattices are complete and equal, every fence counted `parts`, and the ticket word is back at zero.
int emu_beam_in_parts(const emu_scene *s, const FrameParams *f, int W, int H, int parts, float *beam, int nthreads) {
  const SceneView sc = emu_view_of(s, nullptr);
  const size_t n = (size_t)(W / 4 + 1) * (size_t)(H / 4 + 1);
  const float poison = -7.0f;
  std::vector<float> own(n + 4, poison), peer(n + 4, poison);
  unsigned int fences[2 * 8 * 16] = {0}, ticket[2] = {0, 0};
  const int lh = H / 4 + 1, slot = 9;
  simt::g_os_threads = nthreads;
  for (int r = parts - 1; r >= 0; r--) {
    FenceList dst, sig;
    dst.n = 2;
    dst.p[0] = (unsigned int *)own.data();
    dst.p[1] = (unsigned int *)peer.data();
    sig.n = 2;
    sig.p[0] = fences + 8 * slot;
    sig.p[1] = fences + 8 * 16 + 8 * slot;
    if (launch_beam_lattice_rows(sc, *f, own.data(), W, H, r * lh / parts, (r + 1) * lh / parts, dst, sig, ticket, nullptr) != cudaSuccess) return 1;
  }
  if (fences[8 * slot] != (unsigned)parts || fences[8 * 16 + 8 * slot] != (unsigned)parts || ticket[0] != 0u) return 2;
  for (size_t i = 0; i < n; i++)
    if (own[i] == poison || memcmp(&own[i], &peer[i], 4) != 0) return 3;
  return (int)launch_beam_filter(peer.data(), beam, W, H, nullptr);
}

// svo_cast through the product's launch_cast on the emulator.  kernel 0 = grid-stride kernel, 1 = persistent threads with
// warp-level ray fetch; order = optional permutation (the binned order of SVO_OPT_RAY_SORT).
int emu_launch_cast(const emu_scene *s, const void *rays, const uint32_t *order, uint64_t n, void *out, int maxDepth, int kernel, int ctas,
                    int nthreads) {
  const SceneView sc = emu_view_of(s, nullptr);
  unsigned int counter = 0;
  LaunchCfg cfg;
  cfg.fast = false;
  cfg.aux = false;
  cfg.box = false;
  cfg.kernel = 0;
  cfg.stream_kernel = kernel;
  cfg.sm_count = ctas;
  cfg.scene_levels = 8;
  cfg.ctas_per_sm = 1;
  cfg.band_stride = cfg.band_offset = 0;
  cfg.band_ctas = 1;
  cfg.tile_counter = &counter;
  cfg.tile_queue = nullptr;
  cfg.fences.n = 0;
  cfg.split = SplitQueue();
  simt::g_os_threads = kernel == 1 ? 1 : nthreads;
  return (int)launch_cast(cfg, sc, rays, order, n, out, maxDepth, nullptr);
}

}  // extern "C"
