// wavefront_emu.cpp -- TEST INFRASTRUCTURE: svo_wavefront.cu (kernel variant 2: traversal and shading kernels alternating
// over device queues) compiled by g++ and run on the coroutine SIMT emulator, workspace in host memory.
#include "cuda_host_shim.h"
#include "simt_emu.h"

#define SVO_HOST_EMU 1
#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), cudaSuccess)
#define cudaGetLastError() cudaSuccess
#include "../../svo_raytracer_b200/csrc/svo_wavefront.cu"
#undef cudaMemsetAsync
#undef cudaGetLastError

#include <vector>

using namespace svo;

int emu_launch_wavefront(const LaunchCfg &cfg, const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H, int y0, int y1) {
  // ensure_wavefront() of svo_capi.cu
  const uint64_t slots = (uint64_t)((W + 7) / 8) * (uint64_t)((H + 3) / 4) * 32u;
  std::vector<uint4> block((size_t)(18 * slots));
  std::vector<unsigned> counters(2 * (kWaveMaxStages + 1), 0u);
  WaveWorkspace ws;
  uint4 *p = block.data();
  auto take = [&]() { uint4 *r = p; p += slots; return r; };
  for (int q = 0; q < 2; q++) { ws.rayA[q] = take(); ws.rayB[q] = take(); }
  ws.hitA = take();
  ws.hitB = take();
  for (int q = 0; q < 2; q++)
    for (int k = 0; k < 6; k++) ws.state[q][k] = take();
  ws.counters = counters.data();
  ws.slots = slots;
  return (int)launch_render_wavefront(cfg, sc, f, pl, W, H, y0, y1, ws, nullptr);
}
