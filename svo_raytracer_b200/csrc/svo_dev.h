// svo_dev.h -- device-memory and launch helpers shared by the upload-time passes (svo_gpu_build.cu, svo_gpu_transcode.cu).
// The CPU test suite compiles those files with g++ and runs their kernels on the coroutine SIMT emulator
// (tests/hostemu/simt_emu.h, SVO_HOST_EMU): "device" memory is then host memory and the scan is a loop.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>

#ifndef SVO_HOST_EMU
#include <cub/cub.cuh>
#endif

#ifndef SVO_LAUNCH
#ifdef SVO_HOST_EMU
#define SVO_LAUNCH(grid, block, stream, ...) simt::launcher(grid, block, __VA_ARGS__)
#else
#define SVO_LAUNCH(grid, block, stream, ...) __VA_ARGS__<<<grid, block, 0, stream>>>
#endif
#endif

namespace svo {
namespace dev {

#ifdef SVO_HOST_EMU
inline cudaError_t alloc(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline void release(void *p) { free(p); }
inline cudaError_t copy(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) { memcpy(dst, src, bytes); return cudaSuccess; }
inline cudaError_t fill(void *p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
inline cudaError_t sync(cudaStream_t) { return cudaSuccess; }
inline cudaError_t last_error() { return cudaSuccess; }
// in-place exclusive prefix sum over n words; temp_bytes == 0: size query
inline cudaError_t exclusive_scan(void *, size_t &temp_bytes, uint32_t *data, int n, cudaStream_t) {
  if (temp_bytes == 0) { temp_bytes = 1; return cudaSuccess; }
  uint32_t run = 0;
  for (int i = 0; i < n; i++) { const uint32_t v = data[i]; data[i] = run; run += v; }
  return cudaSuccess;
}
#else
inline cudaError_t alloc(void **p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 16); }
inline void release(void *p) { cudaFree(p); }
inline cudaError_t copy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) { return cudaMemcpyAsync(dst, src, bytes, kind, st); }
inline cudaError_t fill(void *p, int v, size_t bytes, cudaStream_t st) { return cudaMemsetAsync(p, v, bytes, st); }
inline cudaError_t sync(cudaStream_t st) { return cudaStreamSynchronize(st); }
inline cudaError_t last_error() { return cudaGetLastError(); }
inline cudaError_t exclusive_scan(void *temp, size_t &temp_bytes, uint32_t *data, int n, cudaStream_t st) {
  return cub::DeviceScan::ExclusiveSum(temp_bytes == 0 ? nullptr : temp, temp_bytes, data, data, n, st);
}
#endif

struct Pool {  // allocations of one pass, freed together
  void *blocks[256];
  int n = 0;
  ~Pool() { for (int i = 0; i < n; i++) release(blocks[i]); }
  template <class T> cudaError_t get(T **p, size_t count) {
    if (n >= 256) return cudaErrorMemoryAllocation;
    void *q = nullptr;
    cudaError_t e = alloc(&q, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) { blocks[n++] = q; *p = (T *)q; }
    return e;
  }
};

}  // namespace dev
}  // namespace svo
