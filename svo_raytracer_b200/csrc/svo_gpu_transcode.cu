// svo_gpu_transcode.cu -- the upload-time transcode (svo_transcode.h) on the device.
//
// Same result, bit for bit, as the host version in svo_transcode.cpp (tests compare the two): a level-synchronous
// breadth-first walk of the reference node stream (src/engine/Octree.java:68-95,119-176) that already sits in HBM
// after the upload copy.  Per level: one thread per node counts the node's descriptor-bearing children, an
// exclusive scan (CUB) turns the counts into slots, a second kernel emits the 8-byte descriptors, the reference
// child-block offsets and the next level's work list at their final positions, and reduces the bounds of
// everything a cast can hit.  ~25 launches for a 13-level tree; the byte-granular record reads are served by L1/L2
// (siblings are contiguous), so the pass is bounded by HBM streaming of the raw stream once per pass.
// Why on the device: svo_upload_range (the engine's per-edit glBufferSubData, Renderer.java:136-146) has to refresh
// the descriptors after every SDF edit; re-walking 2 GB on the host costs seconds, here tens of milliseconds.
#include <cub/cub.cuh>

#include "svo_kernels.h"
#include "svo_transcode.h"

namespace svo {
namespace {

struct Work {  // one node of the current level
  uint32_t off, x, y, z;
};

__device__ __forceinline__ uint32_t rdb(const uint8_t *__restrict__ raw, uint64_t n, uint32_t p) {
  return ((uint64_t)p < n) ? (uint32_t)__ldg(raw + p) : 0u;  // getByte, out of range = 0
}
__device__ __forceinline__ uint32_t rd32(const uint8_t *__restrict__ raw, uint64_t n, uint32_t p) {
  return (rdb(raw, n, p) << 24) | (rdb(raw, n, p + 1u) << 16) | (rdb(raw, n, p + 2u) << 8) | rdb(raw, n, p + 3u);
}
__device__ __forceinline__ uint32_t rd16(const uint8_t *__restrict__ raw, uint64_t n, uint32_t p) {
  return (rdb(raw, n, p) << 8) | rdb(raw, n, p + 1u);
}

// bounds[0..2] = min, [3..5] = max of one CellBox
__device__ __forceinline__ void box_reduce(uint32_t *bounds, const uint32_t lo[3], const uint32_t hi[3], bool any) {
  const unsigned lane = threadIdx.x & 31u;
  const bool warp_any = __any_sync(0xffffffffu, any);
  if (!warp_any) return;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const uint32_t l = __reduce_min_sync(0xffffffffu, any ? lo[a] : 0xFFFFFFFFu);
    const uint32_t h = __reduce_max_sync(0xffffffffu, any ? hi[a] : 0u);
    if (lane == 0) {
      atomicMin(bounds + a, l);
      atomicMax(bounds + 3 + a, h);
    }
  }
}

template <bool EMIT>
__global__ void __launch_bounds__(256) k_transcode_level(const uint8_t *__restrict__ raw, uint64_t nbytes, const Work *__restrict__ cur,
                                                         uint32_t n, int depth, uint32_t *__restrict__ counts,
                                                         const uint32_t *__restrict__ offsets, uint64_t level_base, uint2 *__restrict__ desc,
                                                         uint32_t *__restrict__ refbase, Work *__restrict__ next, uint32_t *__restrict__ leaf_bounds,
                                                         uint32_t *__restrict__ depth_bounds) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t llo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, lhi[3] = {0u, 0u, 0u};
  uint32_t dlo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, dhi[3] = {0u, 0u, 0u};
  bool lany = false, dany = false;
  if (i < n) {
    const Work w = cur[i];
    const uint32_t cp = rd32(raw, nbytes, w.off + 1u), codes = rd16(raw, nbytes, w.off + 5u);
    const uint32_t ref_base = w.off + cp;  // uint wrap-around as in extractChild (svotrace.comp:134)
    uint32_t p = ref_base, nonzero = 0, has_desc = 0, n_next = 0;
    const uint32_t slot0 = EMIT ? offsets[i] : 0u;
    for (uint32_t c = 0; c < 8; c++) {
      const uint32_t code = (codes >> (2u * c)) & 3u;
      const uint32_t size = code == 1u ? 3u : (code == 3u ? 1u : 7u);
      const uint32_t value = rdb(raw, nbytes, p);
      if (value != 0u) {
        nonzero |= 1u << c;
        const uint32_t ccp = code == 0u ? rd32(raw, nbytes, p + 1u) : 0u;
        const uint32_t cx = 2u * w.x + (c & 1u), cy = 2u * w.y + ((c >> 1) & 1u), cz = 2u * w.z + ((c >> 2) & 1u);
        if (EMIT) {
          const int sh = 24 - (depth + 1);
          const uint32_t lo[3] = {cx << sh, cy << sh, cz << sh}, hi[3] = {(cx + 1u) << sh, (cy + 1u) << sh, (cz + 1u) << sh};
          for (int a = 0; a < 3; a++) { dlo[a] = min(dlo[a], lo[a]); dhi[a] = max(dhi[a], hi[a]); }
          dany = true;
          if (ccp == 0u) {  // child.cp == 0: a hit wherever the traversal meets it (svotrace.comp:311)
            for (int a = 0; a < 3; a++) { llo[a] = min(llo[a], lo[a]); lhi[a] = max(lhi[a], hi[a]); }
            lany = true;
          }
        }
        if (ccp != 0u && depth < 22) {  // the traversal may PUSH into it
          has_desc |= 1u << c;
          if (EMIT) {
            Work nw;
            nw.off = p; nw.x = cx; nw.y = cy; nw.z = cz;
            next[slot0 + n_next] = nw;
          }
          n_next++;
        }
      }
      p += size;
    }
    if (EMIT) {
      desc[level_base + i] = make_uint2((uint32_t)(level_base + n + slot0), (codes & 0xFFFFu) | (nonzero << 16) | (has_desc << 24));
      refbase[level_base + i] = ref_base;
    } else {
      counts[i] = n_next;
    }
  }
  if (EMIT) {
    box_reduce(leaf_bounds, llo, lhi, lany);
    box_reduce(depth_bounds + 6 * (depth + 1), dlo, dhi, dany);
  }
}

__global__ void k_init_bounds(uint32_t *b, int nboxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nboxes * 6) b[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;
}

}  // namespace

// Device-side transcode.  `desc`/`refbase` must hold `cap` entries; returns the number written in *ndesc.
// Returns cudaSuccess and sets *overflow if the tree needs more than `cap` descriptors (not a tree, or cap too small).
cudaError_t gpu_transcode(const uint8_t *d_raw, uint64_t nbytes, uint2 *desc, uint32_t *refbase, uint64_t cap, uint64_t *ndesc,
                          uint32_t *nlevels, CellBox *leaf_box, CellBox *depth_box, bool *overflow, cudaStream_t stream) {
  *overflow = false;
  *ndesc = 0;
  *nlevels = 0;
  cudaError_t e;
  Work *buf[2] = {nullptr, nullptr};
  uint32_t *counts = nullptr, *offsets = nullptr, *bounds = nullptr;
  void *scan_tmp = nullptr;
  size_t scan_bytes = 0, work_cap = 0;
  auto cleanup = [&]() {
    cudaFree(buf[0]); cudaFree(buf[1]); cudaFree(counts); cudaFree(offsets); cudaFree(bounds); cudaFree(scan_tmp);
  };
  // work lists and scan buffers for levels of up to n nodes; the `live` entries of buf[keep] survive a regrow
  auto ensure = [&](size_t n, int keep, size_t live) -> cudaError_t {
    if (n <= work_cap) return cudaSuccess;
    cudaError_t s = cudaStreamSynchronize(stream);
    if (s != cudaSuccess) return s;
    Work *old[2] = {buf[0], buf[1]};
    work_cap = n + n / 2 + 1024;
    for (int k = 0; k < 2; k++) {
      Work *nb = nullptr;
      if ((s = cudaMalloc((void **)&nb, work_cap * sizeof(Work))) != cudaSuccess) return s;
      if (k == keep && old[k] && live) cudaMemcpy(nb, old[k], live * sizeof(Work), cudaMemcpyDeviceToDevice);
      cudaFree(old[k]);
      buf[k] = nb;
    }
    cudaFree(counts); cudaFree(offsets); cudaFree(scan_tmp);
    counts = offsets = nullptr;
    scan_tmp = nullptr;
    if ((s = cudaMalloc((void **)&counts, work_cap * sizeof(uint32_t))) != cudaSuccess) return s;
    if ((s = cudaMalloc((void **)&offsets, work_cap * sizeof(uint32_t))) != cudaSuccess) return s;
    scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts, offsets, (int)work_cap);
    return cudaMalloc(&scan_tmp, scan_bytes ? scan_bytes : 16);
  };
  if ((e = cudaMalloc((void **)&bounds, 25 * 6 * sizeof(uint32_t))) != cudaSuccess) { cleanup(); return e; }
  k_init_bounds<<<1, 256, 0, stream>>>(bounds, 25);
  if ((e = ensure(1 << 16, 0, 0)) != cudaSuccess) { cleanup(); return e; }
  const Work root = {0u, 0u, 0u, 0u};
  if ((e = cudaMemcpyAsync(buf[0], &root, sizeof root, cudaMemcpyHostToDevice, stream)) != cudaSuccess) { cleanup(); return e; }
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) { cleanup(); return e; }  // `root` is a stack variable

  uint64_t level_base = 0;
  uint32_t n = 1;
  int cur = 0;
  for (int depth = 0; depth <= 22 && n > 0; depth++) {
    // (8 children per node: from 2^29 nodes on, the 32-bit prefix sum below could wrap -- leave such a level to the host pass)
    if (level_base + n > cap || n >= (1u << 29)) { *overflow = true; break; }
    const unsigned grid = (n + 255) / 256;
    k_transcode_level<false><<<grid, 256, 0, stream>>>(d_raw, nbytes, buf[cur], n, depth, counts, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    size_t sb = scan_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, sb, counts, offsets, (int)n, stream)) != cudaSuccess) break;
    uint32_t last[2] = {0, 0};
    if ((e = cudaMemcpyAsync(&last[0], offsets + (n - 1), 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(&last[1], counts + (n - 1), 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) break;
    const uint64_t total = (uint64_t)last[0] + last[1];
    if (level_base + n + total > cap) { *overflow = true; break; }
    if (total > work_cap) {
      // grow the NEXT list only (the current one is still needed): simplest is to regrow both, preserving contents
      if ((e = ensure(total, cur, n)) != cudaSuccess) break;
      // counts/offsets were reallocated: recompute them for this level
      k_transcode_level<false><<<grid, 256, 0, stream>>>(d_raw, nbytes, buf[cur], n, depth, counts, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
      sb = scan_bytes;
      if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, sb, counts, offsets, (int)n, stream)) != cudaSuccess) break;
    }
    k_transcode_level<true><<<grid, 256, 0, stream>>>(d_raw, nbytes, buf[cur], n, depth, nullptr, offsets, level_base, desc, refbase, buf[cur ^ 1],
                                                      bounds, bounds + 6);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    level_base += n;
    (*nlevels)++;
    n = (uint32_t)total;
    cur ^= 1;
  }
  if (e == cudaSuccess) {
    uint32_t hb[25 * 6];
    e = cudaMemcpyAsync(hb, bounds, sizeof hb, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) {
      auto load = [&](CellBox &b, const uint32_t *p) { for (int a = 0; a < 3; a++) { b.lo[a] = p[a]; b.hi[a] = p[3 + a]; } };
      load(*leaf_box, hb);
      for (int d = 0; d < 24; d++) load(depth_box[d], hb + 6 * (d + 1));
    }
  }
  *ndesc = level_base;
  cleanup();
  return e;
}

}  // namespace svo
