// gpu_build_emu.cpp -- TEST INFRASTRUCTURE: svo_gpu_build.cu (world generation on the device) compiled by g++ and run on
// the coroutine SIMT emulator: the kernels' statements, warp shuffles included, with device memory = host memory and
// the CUB scan replaced by a loop.  tests/test_builder.py compares its stream with svo_build.cpp's byte for byte.
#include "cuda_host_shim.h"
#include "simt_emu.h"

#define SVO_HOST_EMU 1
#include "../../svo_raytracer_b200/csrc/svo_gpu_build.cu"

extern "C" int emu_gpu_build_terrain(const uint16_t *height, const uint8_t *mat, int n, int chunk, uint8_t *out, uint64_t cap,
                                     uint64_t *out_bytes, int nthreads) {
  uint8_t *stream = nullptr;
  uint64_t nbytes = 0, capacity = 0, launches = 0;
  bool unsupported = false;
  simt::g_os_threads = nthreads;
  const cudaError_t e = svo::gpu_build_terrain(height, mat, n, chunk, &stream, &nbytes, &capacity, &unsupported, nullptr, &launches);
  if (e != cudaSuccess) return 1000 + (int)e;
  if (unsupported) return 1;
  *out_bytes = nbytes;
  int rc = 0;
  if (out && cap >= nbytes) memcpy(out, stream, nbytes);
  else if (out) rc = 2;
  free(stream);
  return rc;
}
