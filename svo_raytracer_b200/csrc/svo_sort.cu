// svo_sort.cu -- ray binning for ray streams (svo_cast with SVO_OPT_RAY_SORT, BASELINE configs[3]).
//
// Incoherent secondary rays are ordered by (direction octant, Morton code of the origin in the octree cube) before
// tracing: rays of one octant walk the children in the same order (same octant_mask, svotrace.comp:238-241) and
// rays that start in the same region share the upper-level descent, so a warp's 32 rays agree on PUSH/ADVANCE/POP
// more often and touch fewer distinct descriptors.  Key generation is a hand-written kernel; the 30 / 32-bit
// key/index radix sort is CUB's (library plumbing, like the host-side sort it replaces); results are written back
// to the caller's order, so the output is unchanged bit for bit.
#include <cub/cub.cuh>

#include "svo_kernels.h"

namespace svo {

struct RayRecS { float ox, oy, oz, dx, dy, dz; };

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 9 bits -> every third bit
  v &= 0x1FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// mode 1: key = octant (3) | origin Morton code on a 512^3 grid (27).
// mode 2: key = octant (3) | origin Morton code on a 64^3 grid (18) | direction bin inside the octant (6: |d| / (|dx|+|dy|+|dz|)
//         in 8 x 8 cells) | the next 5 Morton bits of the origin.  Rays that start in the same 64^3 cell are ordered by direction
//         before position: 32 neighbouring rays then agree on PUSH / ADVANCE / POP for longer after they leave the cell.  SIMT
//         model of the loop on dense-origin streams (tools/simt_model.py machinery): 118.6 M -> 100.3 M issue slots (-15 %).
__global__ void __launch_bounds__(256) k_ray_keys(const RayRecS *__restrict__ rays, uint64_t n, uint32_t *__restrict__ keys,
                                                  uint32_t *__restrict__ idx, int mode) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RayRecS r = rays[i];
  const uint32_t oct = (r.dx > 0.0f ? 1u : 0u) | (r.dy > 0.0f ? 2u : 0u) | (r.dz > 0.0f ? 4u : 0u);
  // origin quantised to a 512^3 grid over the cube [1,2]^3 (clamped: origins may lie outside, NaN -> 0)
  const float qx = fminf(fmaxf((r.ox - 1.0f) * 512.0f, 0.0f), 511.0f);
  const float qy = fminf(fmaxf((r.oy - 1.0f) * 512.0f, 0.0f), 511.0f);
  const float qz = fminf(fmaxf((r.oz - 1.0f) * 512.0f, 0.0f), 511.0f);
  const uint32_t m = spread3((uint32_t)qx) | (spread3((uint32_t)qy) << 1) | (spread3((uint32_t)qz) << 2);
  if (mode == 2) {
    const float ax = fabsf(r.dx), ay = fabsf(r.dy), az = fabsf(r.dz);
    const float sum = ax + ay + az;
    const float inv = sum > 0.0f && sum < 3.0e38f ? 8.0f / sum : 0.0f;  // (zero, NaN and infinite directions: bin 0)
    const uint32_t u = (uint32_t)fminf(fmaxf(ax * inv, 0.0f), 7.0f), v = (uint32_t)fminf(fmaxf(ay * inv, 0.0f), 7.0f);
    keys[i] = (oct << 29) | ((m >> 9) << 11) | (((u << 3) | v) << 5) | ((m >> 4) & 31u);
  } else {
    keys[i] = (oct << 27) | m;
  }
  idx[i] = (uint32_t)i;
}

size_t ray_sort_temp_bytes(uint64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                  (uint32_t *)nullptr, (int)n, 0, 32);
  return bytes;
}

// order_out[i] = index of the i-th ray in sorted order.  keys/idx/keys_alt: n words each; temp: ray_sort_temp_bytes(n).
cudaError_t launch_ray_sort(const void *d_rays, uint64_t n, uint32_t *keys, uint32_t *keys_alt, uint32_t *idx, uint32_t *order_out,
                            void *temp, size_t temp_bytes, int mode, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  k_ray_keys<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const RayRecS *)d_rays, n, keys, idx, mode);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_alt, idx, order_out, (int)n, 0, mode == 2 ? 32 : 30, stream);
}

}  // namespace svo
