// svo_gpu_build.h -- world generation on the device (svo_gpu_build.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace svo {
cudaError_t gpu_build_terrain(const uint16_t *height, const uint8_t *mat, int n, int chunk, uint8_t **stream, uint64_t *nbytes,
                              uint64_t *capacity, bool *unsupported, cudaStream_t st, uint64_t *launches);
}
