// svo_gpu_transcode.cu -- the upload-time transcode (svo_transcode.h) on the device, whole (svo_upload) and
// incremental (svo_upload_range).
//
// Whole stream: same result, bit for bit, as the host version in svo_transcode.cpp (tests compare the two): a
// level-synchronous breadth-first walk of the reference node stream (src/engine/Octree.java:68-95,119-176) that already
// sits in HBM after the upload copy.  Per level: one thread per node counts the node's descriptor-bearing children, an
// exclusive scan turns the counts into slots, a second kernel emits the 8-byte descriptors, the reference child-block
// offsets, the per-node meta record (own record offset, parent index) and the next level's work list at their final
// positions, and reduces the bounds of everything a cast can hit.  ~25 launches for a 13-level tree.
//
// Incremental (gpu_patch): the engine's SDF brush rewrites a few records in place and appends new ones
// (Octree.useSDFBrush / subdivideNode, Octree.java:700-885) and pushes the touched byte ranges with two glBufferSubData
// calls per edit (Main.java:349-350, Renderer.java:136-146).  Re-walking 2 GB for that is ~0.1 s; here
//   1. k_diff_apply compares the new bytes with the old ones while storing them: a bitmap of the bytes that really changed;
//   2. k_mark_dirty tests every node's own record and child block against the bitmap (two compares reject almost all);
//   3. k_select_roots keeps the dirty nodes without a dirty ancestor and finds their cell by climbing the parent links;
//   4. the same level kernels re-walk only the subtrees under those roots: the roots' descriptors are rewritten in place,
//      everything below is appended behind the existing array (old descendants become unreachable garbage until the
//      next whole transcode compacts them).
// A 1 KB in-place edit of the 8192^3 world costs one pass over the per-node meta array (~0.7 GB) plus a handful of
// tiny launches.
#include "svo_dev.h"
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

// bounds[0..2] = min, [3..5] = max of one CellBox.  WARP: all lanes of the warp reduce into the same box first.
template <bool WARP>
__device__ __forceinline__ void box_reduce(uint32_t *bounds, const uint32_t lo[3], const uint32_t hi[3], bool any) {
  if (!WARP) {
    if (any)
      for (int a = 0; a < 3; a++) { atomicMin(bounds + a, lo[a]); atomicMax(bounds + 3 + a, hi[a]); }
    return;
  }
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

// One level of the walk.  EMIT false: counts[i] = descriptor-bearing children of node i.  EMIT true: node i's descriptor
// goes to out_index[i] (patch: a root rewritten in place) or level_base + i; its children's descriptors will live at
// next_base + offsets[i] ...; meta (own record offset, parent index) is written for those children here.
// VARDEPTH: nodes of one launch sit at different depths (depth_of[i]); otherwise all at `depth`.
template <bool EMIT, bool VARDEPTH>
__global__ void __launch_bounds__(256) k_transcode_level(const uint8_t *__restrict__ raw, uint64_t nbytes, const Work *__restrict__ cur,
                                                         uint32_t n, int depth_uniform, const uint8_t *__restrict__ depth_of,
                                                         uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets,
                                                         uint64_t level_base, const uint32_t *__restrict__ out_index, uint64_t next_base,
                                                         uint2 *__restrict__ desc, uint32_t *__restrict__ refbase, uint2 *__restrict__ meta,
                                                         Work *__restrict__ next, uint8_t *__restrict__ next_depth,
                                                         uint32_t *__restrict__ leaf_bounds, uint32_t *__restrict__ depth_bounds) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t llo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, lhi[3] = {0u, 0u, 0u};
  uint32_t dlo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, dhi[3] = {0u, 0u, 0u};
  bool lany = false, dany = false;
  int depth = depth_uniform;
  if (i < n) {
    if (VARDEPTH) depth = (int)depth_of[i];
    const Work w = cur[i];
    const uint32_t cp = rd32(raw, nbytes, w.off + 1u), codes = rd16(raw, nbytes, w.off + 5u);
    const uint32_t ref_base = w.off + cp;  // uint wrap-around as in extractChild (svotrace.comp:134)
    uint32_t p = ref_base, nonzero = 0, has_desc = 0, n_next = 0;
    const uint32_t slot0 = EMIT ? offsets[i] : 0u;
    const uint32_t my = EMIT ? (out_index ? out_index[i] : (uint32_t)(level_base + i)) : 0u;
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
            if (next_depth) next_depth[slot0 + n_next] = (uint8_t)(depth + 1);
            if (meta) meta[next_base + slot0 + n_next] = make_uint2(p, my);
          }
          n_next++;
        }
      }
      p += size;
    }
    if (EMIT) {
      desc[my] = make_uint2((uint32_t)(next_base + slot0), (codes & 0xFFFFu) | (nonzero << 16) | (has_desc << 24));
      refbase[my] = ref_base;
    } else {
      counts[i] = n_next;
    }
  }
  if (EMIT) {
    box_reduce<!VARDEPTH>(leaf_bounds, llo, lhi, lany);
    box_reduce<!VARDEPTH>(depth_bounds + 6 * (depth + 1), dlo, dhi, dany);
  }
}

__global__ void k_init_bounds(uint32_t *b, int nboxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nboxes * 6) b[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;
}

// ---- incremental ------------------------------------------------------------------------------------------------
// bit j of bitmap = byte (start + j) of the stream changed.  Thread t handles bytes [start + 8t, start + 8t + 8).
__global__ void __launch_bounds__(256) k_diff_apply(uint8_t *__restrict__ raw, const uint8_t *__restrict__ fresh, uint64_t start, uint64_t end,
                                                    uint64_t old_nbytes, uint8_t *__restrict__ bitmap, unsigned long long *__restrict__ span) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t b0 = start + 8u * t;
  if (b0 >= end) return;
  uint32_t bits = 0;
  for (uint32_t k = 0; k < 8u && b0 + k < end; k++) {
    const uint8_t nv = fresh[b0 - start + k];
    const bool was = b0 + k < old_nbytes;  // bytes beyond the old end read as zero (rdb)
    const uint8_t ov = was ? raw[b0 + k] : (uint8_t)0;
    if (nv != ov) bits |= 1u << k;
    raw[b0 + k] = nv;
  }
  bitmap[t] = (uint8_t)bits;
  if (bits) {
    atomicMin(span + 0, (unsigned long long)(b0 + (uint32_t)(__ffs((int)bits) - 1)));
    atomicMax(span + 1, (unsigned long long)(b0 + (uint32_t)(31 - __clz((int)bits)) + 1u));
  }
}

// any changed byte in [a, b)?  (b - a <= 64)
__device__ __forceinline__ bool range_changed(const uint8_t *__restrict__ bitmap, uint64_t start, uint64_t end, uint64_t lo, uint64_t hi,
                                              uint64_t a, uint64_t b) {
  if (a < lo) a = lo;
  if (b > hi) b = hi;
  if (a >= b) return false;
  (void)end;
  for (uint64_t byte = (a - start) >> 3; byte <= (b - 1 - start) >> 3; byte++) {
    uint32_t m = bitmap[byte];
    if (!m) continue;
    const uint64_t base = start + (byte << 3);
    for (uint32_t k = 0; k < 8u; k++)
      if (((m >> k) & 1u) && base + k >= a && base + k < b) return true;
  }
  return false;
}

// A node's descriptor depends on its own record's child pointer and type mask, and on the eight records of its child
// block (values, child pointers).  Dirty nodes go to `list`; flag[i] = 1.
__global__ void __launch_bounds__(256) k_mark_dirty(const uint2 *__restrict__ desc, const uint32_t *__restrict__ refbase, const uint2 *__restrict__ meta,
                                                    uint32_t ndesc, const uint8_t *__restrict__ bitmap, uint64_t start, uint64_t end, uint64_t lo,
                                                    uint64_t hi, uint8_t *__restrict__ flag, uint32_t *__restrict__ list, uint32_t list_cap,
                                                    uint32_t *__restrict__ count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ndesc) return;
  const uint64_t r = meta[i].x, b = refbase[i];
  // own record bytes 1..6; child block: at most 56 bytes
  const bool near_own = r + 7u > lo && r + 1u < hi, near_block = b + 56u > lo && b < hi;
  if (!near_own && !near_block) return;
  bool dirty = near_own && range_changed(bitmap, start, end, lo, hi, r + 1u, r + 7u);
  if (!dirty && near_block) {
    const uint32_t codes = desc[i].y & 0xFFFFu;  // the OLD type codes give the block's old extent; new codes mean the own record changed
    uint32_t size = 0;
    for (uint32_t c = 0; c < 8u; c++) {
      const uint32_t code = (codes >> (2u * c)) & 3u;
      size += code == 1u ? 3u : (code == 3u ? 1u : 7u);
    }
    dirty = range_changed(bitmap, start, end, lo, hi, b, b + size);
  }
  if (dirty) {
    flag[i] = 1;
    const uint32_t k = atomicAdd(count, 1u);
    if (k < list_cap) list[k] = i;
  }
}

// Dirty nodes without a dirty ancestor become the roots of the re-walk; their cell follows from the child slots on the way up.
__global__ void __launch_bounds__(128) k_select_roots(const uint2 *__restrict__ desc, const uint2 *__restrict__ meta, const uint8_t *__restrict__ flag,
                                                      const uint32_t *__restrict__ list, uint32_t nlist, Work *__restrict__ roots,
                                                      uint8_t *__restrict__ root_depth, uint32_t *__restrict__ root_index, uint32_t *__restrict__ nroots) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nlist) return;
  const uint32_t i = list[t];
  uint32_t slots[24];
  int depth = 0;
  uint32_t j = i;
  while (j != 0u) {
    const uint32_t p = meta[j].y;
    if (p == 0xFFFFFFFFu || depth >= 23) return;  // an orphan of an earlier patch (or a corrupt chain): unreachable, nothing to do
    if (flag[p]) return;                           // a dirty ancestor re-walks this node anyway
    const uint2 pd = desc[p];
    const uint32_t has = pd.y >> 24, k = j - pd.x;  // j is the k-th descriptor-bearing child of p
    uint32_t slot = 8u;
    if (k < 8u) {
      uint32_t seen = 0;
      for (uint32_t c = 0; c < 8u; c++)
        if ((has >> c) & 1u) {
          if (seen == k) { slot = c; break; }
          seen++;
        }
    }
    if (slot == 8u) return;  // p no longer points at j: j is garbage
    slots[depth++] = slot;
    j = p;
  }
  uint32_t x = 0, y = 0, z = 0;
  for (int d = depth - 1; d >= 0; d--) {
    x = 2u * x + (slots[d] & 1u);
    y = 2u * y + ((slots[d] >> 1) & 1u);
    z = 2u * z + ((slots[d] >> 2) & 1u);
  }
  const uint32_t k = atomicAdd(nroots, 1u);
  Work w;
  w.off = meta[i].x; w.x = x; w.y = y; w.z = z;
  roots[k] = w;
  root_depth[k] = (uint8_t)depth;
  root_index[k] = i;
}

__global__ void k_clear_flags(uint8_t *__restrict__ flag, const uint32_t *__restrict__ list, uint32_t n) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) flag[list[t]] = 0;
}

void load_boxes(const uint32_t *hb, CellBox *leaf_box, CellBox *depth_box) {
  auto load = [&](CellBox &b, const uint32_t *p) { for (int a = 0; a < 3; a++) { b.lo[a] = p[a]; b.hi[a] = p[3 + a]; } };
  load(*leaf_box, hb);
  for (int d = 0; d < 24; d++) load(depth_box[d], hb + 6 * (d + 1));
}

// Scratch of the incremental path: a caller-owned arena (no cudaMalloc per edit: each costs more than the whole patch),
// spilling into fresh allocations only for edits that outgrow it.
struct Scratch {
  uint8_t *base;
  size_t cap, used;
  dev::Pool pool;
  template <class T> cudaError_t get(T **p, size_t count) {
    const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
    if (base && used + bytes <= cap) {
      *p = (T *)(base + used);
      used += bytes;
      return cudaSuccess;
    }
    return pool.get(p, count);
  }
};

#define GT_CUDA(call)                   \
  do {                                  \
    cudaError_t e_ = (call);            \
    if (e_ != cudaSuccess) return e_;   \
  } while (0)

}  // namespace

// Device-side transcode.  `desc` / `refbase` / `meta` (may be null) must hold `cap` entries; returns the number written in *ndesc.
// Returns cudaSuccess and sets *overflow if the tree needs more than `cap` descriptors (not a tree, or cap too small).
cudaError_t gpu_transcode(const uint8_t *d_raw, uint64_t nbytes, uint2 *desc, uint32_t *refbase, uint2 *meta, uint64_t cap, uint64_t *ndesc,
                          uint32_t *nlevels, CellBox *leaf_box, CellBox *depth_box, bool *overflow, cudaStream_t stream) {
  *overflow = false;
  *ndesc = 0;
  *nlevels = 0;
  dev::Pool pool;
  uint32_t *bounds = nullptr;
  GT_CUDA(pool.get(&bounds, 25 * 6));
  SVO_LAUNCH(1, 256, stream, k_init_bounds)(bounds, 25);
  // two work lists and a count / offset array, regrown (by fresh allocations; the pool frees the old ones at the end) when a level outgrows them
  size_t work_cap = 1 << 16, scan_bytes = 0;
  Work *buf[2] = {nullptr, nullptr};
  uint32_t *offsets = nullptr;
  void *scan_tmp = nullptr;
  auto grow = [&](size_t n) -> cudaError_t {
    work_cap = n + n / 2 + 1024;
    GT_CUDA(pool.get(&buf[0], work_cap));
    GT_CUDA(pool.get(&buf[1], work_cap));
    GT_CUDA(pool.get(&offsets, work_cap + 1));
    scan_bytes = 0;
    GT_CUDA(dev::exclusive_scan(nullptr, scan_bytes, offsets, (int)work_cap + 1, stream));
    return pool.get((uint8_t **)&scan_tmp, scan_bytes);
  };
  GT_CUDA(grow(work_cap));
  const Work root = {0u, 0u, 0u, 0u};
  const uint2 root_meta = make_uint2(0u, 0xFFFFFFFFu);
  GT_CUDA(dev::copy(buf[0], &root, sizeof root, cudaMemcpyHostToDevice, stream));
  if (meta && cap > 0) GT_CUDA(dev::copy(meta, &root_meta, sizeof root_meta, cudaMemcpyHostToDevice, stream));
  GT_CUDA(dev::sync(stream));  // `root` is a stack variable

  uint64_t level_base = 0;
  uint32_t n = 1;
  Work *cur = buf[0], *nxt = buf[1];
  for (int depth = 0; depth <= 22 && n > 0; depth++) {
    // (8 children per node: from 2^29 nodes on, the 32-bit prefix sum below could wrap -- leave such a level to the host pass)
    if (level_base + n > cap || n >= (1u << 29)) { *overflow = true; break; }
    const unsigned grid = (n + 255) / 256;
    SVO_LAUNCH(grid, 256, stream, k_transcode_level<false, false>)(d_raw, nbytes, cur, n, depth, nullptr, offsets, nullptr, 0, nullptr, 0, nullptr,
                                                                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    GT_CUDA(dev::fill(offsets + n, 0, sizeof(uint32_t), stream));
    size_t sb = scan_bytes;
    GT_CUDA(dev::exclusive_scan(scan_tmp, sb, offsets, (int)n + 1, stream));  // in place; entry n becomes the total
    uint32_t total32 = 0;
    GT_CUDA(dev::copy(&total32, offsets + n, 4, cudaMemcpyDeviceToHost, stream));
    GT_CUDA(dev::sync(stream));
    const uint64_t total = total32;
    if (level_base + n + total > cap) { *overflow = true; break; }
    if (total > work_cap) {  // the next list must grow; the current list and its offsets stay where they are
      Work *keep = cur;
      uint32_t *keep_off = offsets;
      GT_CUDA(grow(total));
      cur = keep;
      nxt = buf[1];
      // `offsets` was replaced: the emit pass below still reads the old array
      SVO_LAUNCH(grid, 256, stream, k_transcode_level<true, false>)(d_raw, nbytes, cur, n, depth, nullptr, nullptr, keep_off, level_base, nullptr,
                                                                    level_base + n, desc, refbase, meta, nxt, nullptr, bounds, bounds + 6);
      // from now on both lists are the new, larger ones
      cur = nxt;
      nxt = buf[0];
    } else {
      SVO_LAUNCH(grid, 256, stream, k_transcode_level<true, false>)(d_raw, nbytes, cur, n, depth, nullptr, nullptr, offsets, level_base, nullptr,
                                                                    level_base + n, desc, refbase, meta, nxt, nullptr, bounds, bounds + 6);
      Work *t = cur;
      cur = nxt;
      nxt = t;
    }
    GT_CUDA(dev::last_error());
    level_base += n;
    (*nlevels)++;
    n = (uint32_t)total;
  }
  uint32_t hb[25 * 6];
  GT_CUDA(dev::copy(hb, bounds, sizeof hb, cudaMemcpyDeviceToHost, stream));
  GT_CUDA(dev::sync(stream));
  load_boxes(hb, leaf_box, depth_box);
  *ndesc = level_base;
  return cudaSuccess;
}

// Stores fresh[0, end - start) at d_raw[start, end) and records which bytes changed: bitmap has (end - start + 7) / 8 bytes,
// span[0..1] = [first changed byte, last changed byte + 1) (span[0] > span[1]: nothing changed).
cudaError_t gpu_diff_apply(uint8_t *d_raw, const uint8_t *d_fresh, uint64_t start, uint64_t end, uint64_t old_nbytes, uint8_t *d_bitmap,
                           uint64_t span[2], void *arena, size_t arena_bytes, cudaStream_t stream) {
  Scratch pool = {(uint8_t *)arena, arena_bytes, 0, {}};
  unsigned long long *d_span = nullptr;
  GT_CUDA(pool.get(&d_span, 2));
  const unsigned long long init[2] = {~0ull, 0ull};
  GT_CUDA(dev::copy(d_span, init, sizeof init, cudaMemcpyHostToDevice, stream));
  const uint64_t threads = (end - start + 7) / 8;
  SVO_LAUNCH((unsigned)((threads + 255) / 256), 256, stream, k_diff_apply)(d_raw, d_fresh, start, end, old_nbytes, d_bitmap, d_span);
  unsigned long long h[2];
  GT_CUDA(dev::copy(h, d_span, sizeof h, cudaMemcpyDeviceToHost, stream));
  GT_CUDA(dev::sync(stream));
  span[0] = h[0];
  span[1] = h[1];
  return cudaSuccess;
}

// Incremental transcode after gpu_diff_apply.  A: the scene's arrays (cap entries, *ndesc in use); flag: cap bytes, all zero
// (left all zero).  *fallback = true: the edit is too wide for the incremental path (or the arrays are full): run the whole
// transcode.  stats[0] dirty nodes, [1] roots, [2] descriptors appended.
cudaError_t gpu_patch(const uint8_t *d_raw, uint64_t nbytes, const uint8_t *d_bitmap, uint64_t start, uint64_t end, const uint64_t span[2],
                      uint2 *desc, uint32_t *refbase, uint2 *meta, uint8_t *flag, uint64_t cap, uint64_t *ndesc, CellBox *leaf_box,
                      CellBox *depth_box, bool *fallback, uint64_t stats[3], void *arena, size_t arena_bytes, cudaStream_t stream) {
  *fallback = false;
  stats[0] = stats[1] = stats[2] = 0;
  const uint32_t nd = (uint32_t)*ndesc;
  if (span[0] >= span[1] || nd == 0) return cudaSuccess;
  Scratch pool = {(uint8_t *)arena, arena_bytes, 0, {}};
  const uint32_t list_cap = 1u << 20;
  uint32_t *list = nullptr, *counters = nullptr, *root_index = nullptr, *bounds = nullptr;
  Work *roots = nullptr;
  uint8_t *root_depth = nullptr;
  GT_CUDA(pool.get(&list, list_cap));
  GT_CUDA(pool.get(&counters, 4));
  GT_CUDA(dev::fill(counters, 0, 4 * sizeof(uint32_t), stream));
  SVO_LAUNCH((nd + 255) / 256, 256, stream, k_mark_dirty)(desc, refbase, meta, nd, d_bitmap, start, end, span[0], span[1], flag, list, list_cap, counters);
  uint32_t ndirty = 0;
  GT_CUDA(dev::copy(&ndirty, counters, 4, cudaMemcpyDeviceToHost, stream));
  GT_CUDA(dev::sync(stream));
  stats[0] = ndirty;
  if (ndirty == 0) return cudaSuccess;
  if (ndirty > list_cap) {  // flags of the nodes that did not fit the list cannot be cleared one by one
    GT_CUDA(dev::fill(flag, 0, cap, stream));
    *fallback = true;
    return cudaSuccess;
  }
  GT_CUDA(pool.get(&roots, ndirty));
  GT_CUDA(pool.get(&root_depth, ndirty));
  GT_CUDA(pool.get(&root_index, ndirty));
  SVO_LAUNCH((ndirty + 127) / 128, 128, stream, k_select_roots)(desc, meta, flag, list, ndirty, roots, root_depth, root_index, counters + 1);
  SVO_LAUNCH((ndirty + 255) / 256, 256, stream, k_clear_flags)(flag, list, ndirty);
  uint32_t nroots = 0;
  GT_CUDA(dev::copy(&nroots, counters + 1, 4, cudaMemcpyDeviceToHost, stream));
  GT_CUDA(dev::sync(stream));
  stats[1] = nroots;
  if (nroots == 0) return cudaSuccess;

  // the content boxes only ever grow here (a whole transcode recomputes them)
  GT_CUDA(pool.get(&bounds, 25 * 6));
  uint32_t hb[25 * 6];
  auto store = [&](const CellBox &b, uint32_t *p) { for (int a = 0; a < 3; a++) { p[a] = b.lo[a]; p[3 + a] = b.hi[a]; } };
  store(*leaf_box, hb);
  for (int d = 0; d < 24; d++) store(depth_box[d], hb + 6 * (d + 1));
  GT_CUDA(dev::copy(bounds, hb, sizeof hb, cudaMemcpyHostToDevice, stream));

  // level 0 of the re-walk = the roots, rewritten in place; deeper levels are appended at the tail
  uint64_t tail = nd;
  Work *cur = roots;
  uint8_t *cur_depth = root_depth;
  const uint32_t *out_index = root_index;
  uint32_t n = nroots;
  for (int level = 0; level < 24 && n > 0; level++) {
    uint32_t *offsets = nullptr;
    void *scan_tmp = nullptr;
    size_t scan_bytes = 0;
    GT_CUDA(pool.get(&offsets, (size_t)n + 1));
    GT_CUDA(dev::exclusive_scan(nullptr, scan_bytes, offsets, (int)n + 1, stream));
    GT_CUDA(pool.get((uint8_t **)&scan_tmp, scan_bytes));
    const unsigned grid = (n + 255) / 256;
    SVO_LAUNCH(grid, 256, stream, k_transcode_level<false, true>)(d_raw, nbytes, cur, n, 0, cur_depth, offsets, nullptr, 0, nullptr, 0, nullptr, nullptr,
                                                                  nullptr, nullptr, nullptr, nullptr, nullptr);
    GT_CUDA(dev::fill(offsets + n, 0, sizeof(uint32_t), stream));
    GT_CUDA(dev::exclusive_scan(scan_tmp, scan_bytes, offsets, (int)n + 1, stream));
    uint32_t total = 0;
    GT_CUDA(dev::copy(&total, offsets + n, 4, cudaMemcpyDeviceToHost, stream));
    GT_CUDA(dev::sync(stream));
    const uint64_t level_base = out_index ? 0 : tail;       // where this level's own descriptors go
    const uint64_t next_base = out_index ? tail : tail + n;  // where its children's go
    if (next_base + total > cap) { *fallback = true; return cudaSuccess; }
    Work *nxt = nullptr;
    uint8_t *nxt_depth = nullptr;
    GT_CUDA(pool.get(&nxt, total));
    GT_CUDA(pool.get(&nxt_depth, total));
    SVO_LAUNCH(grid, 256, stream, k_transcode_level<true, true>)(d_raw, nbytes, cur, n, 0, cur_depth, nullptr, offsets, level_base, out_index, next_base,
                                                                 desc, refbase, meta, nxt, nxt_depth, bounds, bounds + 6);
    GT_CUDA(dev::last_error());
    if (!out_index) tail += n;
    out_index = nullptr;
    cur = nxt;
    cur_depth = nxt_depth;
    n = total;
    if (pool.pool.n > 240) { *fallback = true; return cudaSuccess; }  // (cannot happen: at most 24 levels x 4 blocks)
  }
  GT_CUDA(dev::copy(hb, bounds, sizeof hb, cudaMemcpyDeviceToHost, stream));
  GT_CUDA(dev::sync(stream));
  load_boxes(hb, leaf_box, depth_box);
  stats[2] = tail - nd;
  *ndesc = tail;
  return cudaSuccess;
}

}  // namespace svo
