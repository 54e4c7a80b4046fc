// svo_gpu_build.cu -- heightmap world generation ON THE DEVICE, byte for byte the stream svo_build.cpp writes.
//
// What: the node stream of Octree.constructCompleteOctree for a heightmap world (reference src/engine/Octree.java:192-353:
// root + fillEmptyChildren levels, then per chunk eight sub-octrees built by constructInnerOctree :511-608 with
// genSurfaceNormal :620-649 and checkBigNodeExposed :651-670, spliced in :317-337), voxel rule of
// src/shaders/chunkgen-heightmap.comp:13-31.  The reference voxelises every 1024^3 chunk into a 1 GiB image on the GPU,
// reads it back and walks it with 8 Java threads; svo_build.cpp replaced the voxels by min/max pyramids of the height
// and material maps (12-32 s for 8192^3 on the host).  Here the same classification runs as kernels.
//
// How: the recursive, depth-first builder becomes three level-synchronous sweeps over the tree.
//   A  top-down   one thread per (node, child): classify the child cube (homogeneous? value? type) from the pyramids,
//                 assemble the node's type mask / child values / block size by warp shuffles, scan the number of
//                 children that recurse (CUB) and write the next level's node list.
//   B  bottom-up  subtree bytes S(node) = size of its 8-record child block + sum of S over the recursing children:
//                 what constructInnerOctree appends between entering and leaving the node.
//   C  top-down   the byte offset where each node's child block starts (parent's block + the S of earlier siblings:
//                 depth-first order) and where its own record lies; then every node writes its child records (value,
//                 packed normal of surface leaves) and patches child pointer + type mask into its own record.
// The few hundred nodes above the chunks (fill levels, chunk children, the 7-byte gap after every spliced sub-octree,
// Octree.java:336) are laid out on the host from the per-unit sizes.  The stream never leaves HBM: svo_build_terrain_device
// hands it to the same transcode svo_upload runs.
#include <cstdio>
#include <vector>

#include "../../include/svo_b200.h"
#include "svo_dev.h"
#include "svo_gpu_build.h"

namespace svo {
namespace {

using dev::Pool;
inline cudaError_t dev_alloc(void **p, size_t bytes) { return dev::alloc(p, bytes); }
inline void dev_free(void *p) { dev::release(p); }
inline cudaError_t dev_copy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) { return dev::copy(dst, src, bytes, kind, st); }
inline cudaError_t dev_memset(void *p, int v, size_t bytes, cudaStream_t st) { return dev::fill(p, v, bytes, st); }
inline cudaError_t dev_sync(cudaStream_t st) { return dev::sync(st); }
inline cudaError_t dev_last_error() { return dev::last_error(); }
inline cudaError_t exclusive_scan(void *temp, size_t &temp_bytes, uint32_t *data, int n, cudaStream_t st) { return dev::exclusive_scan(temp, temp_bytes, data, n, st); }

enum { T_INTERIOR = 0, T_SURFACE = 1, T_SUBDIV = 2, T_NONSURF = 3 };

struct Pyr {  // min/max pyramids of the column heights and materials; level L has (n >> L)^2 entries
  const uint16_t *hmin[16], *hmax[16];
  const uint8_t *mmin[16], *mmax[16];
  int n, chunk;
};

struct Level {  // one octree level of the sub-octrees (SoA, one entry per node on which constructInnerOctree runs)
  uint32_t *key0, *key1;  // px | py << 16,  pz | unit << 16   (chunk-local origin of the node's cube; unit = chunk * 8 + i)
  uint32_t *packed;       // type codes [0:16) | child-block bytes [16:24) | mask of recursing children [24:32)
  uint2 *vals;            // the 8 child values
  uint32_t *first;        // index of the first recursing child in the next level (exclusive scan of the counts)
  uint32_t *S;            // bytes constructInnerOctree appends for this node
  uint32_t *off;          // stream offset of the node's child block
  uint32_t *recpos;       // stream offset of the node's own record
  uint32_t n;
};

__device__ __forceinline__ int col_h(const Pyr &P, int wx, int wz) { return P.hmin[0][(size_t)wz * (size_t)P.n + (size_t)wx]; }
__device__ __forceinline__ int col_m(const Pyr &P, int wx, int wz) { return P.mmin[0][(size_t)wz * (size_t)P.n + (size_t)wx]; }
// chunkgen-heightmap.comp:22-28
__device__ __forceinline__ uint32_t voxel(const Pyr &P, int wx, int wy, int wz) {
  const int hs = col_h(P, wx, wz);
  if (wy > hs) return 0u;
  return (hs - wy <= 4) ? (uint32_t)col_m(P, wx, wz) : 1u;
}

// The scan of Octree.java:527-555 for the cube (wx0, wy0, wz0, s), s >= 2: is it homogeneous, and its `value`
// (same case analysis as Unit::classify in svo_build.cpp).
__device__ void classify(const Pyr &P, int wx0, int wy0, int wz0, int s, int L, bool &leaf, uint32_t &value) {
  const size_t pi = (size_t)(wz0 >> L) * (size_t)(P.n >> L) + (size_t)(wx0 >> L);
  const int hmn = P.hmin[L][pi], hmx = P.hmax[L][pi];
  const int y1 = wy0 + s - 1;
  const uint32_t first = voxel(P, wx0, wy0, wz0);
  if (wy0 > hmx) { leaf = true; value = 0u; return; }  // all air
  if (y1 > hmn) {                                       // air above the lowest column, rock in the highest
    leaf = false;
    if (first != 0u) { value = first; return; }
    // first == 0: `first` becomes the first non-zero sample in scan order z, y, x: the first column (z-major) that reaches wy0
    for (int z = wz0; z < wz0 + s; z++)
      for (int x = wx0; x < wx0 + s; x++) {
        const int hs = col_h(P, x, z);
        if (hs >= wy0) { value = (hs - wy0 <= 4) ? (uint32_t)col_m(P, x, z) : 1u; return; }
      }
    value = 0u;  // unreachable (hmx >= wy0)
    return;
  }
  value = first;  // every voxel is solid
  if (y1 < hmn - 4) { leaf = true; return; }  // below every material band: all 1
  if (P.mmin[L][pi] == 1 && P.mmax[L][pi] == 1) { leaf = true; return; }
  for (int z = wz0; z < wz0 + s; z++)
    for (int x = wx0; x < wx0 + s; x++) {
      const int hs = col_h(P, x, z);
      const bool band = hs - 4 <= y1;   // some y in range with hs - y <= 4 -> material
      const bool deep = wy0 <= hs - 5;  // some y in range with hs - y >= 5 -> 1
      if ((band && (uint32_t)col_m(P, x, z) != first) || (deep && first != 1u)) { leaf = false; return; }
    }
  leaf = true;
}

// Octree.java:620-649 (neighbours outside the chunk are skipped)
__device__ bool surface_normal(const Pyr &P, int ox, int oy, int oz, int cx, int cy, int cz, uint32_t &packed) {
  bool exposed = false;
  int nx = 0, ny = 0, nz = 0;
  for (int i = cx - 1; i <= cx + 1; i++) {
    if (i < 0 || i >= P.chunk) continue;
    for (int k = cz - 1; k <= cz + 1; k++) {
      if (k < 0 || k >= P.chunk) continue;
      const int hs = col_h(P, ox + i, oz + k);
      for (int j = cy - 1; j <= cy + 1; j++) {
        if (j < 0 || j >= P.chunk) continue;
        if (oy + j > hs) { exposed = true; nx += i - cx; ny += j - cy; nz += k - cz; }  // voxel == 0 iff above the column
      }
    }
  }
  nx = nx / 2 + 5; ny = ny / 2 + 5; nz = nz / 2 + 5;
  packed = (uint32_t)(int)(short)(nx + ny * 10 + nz * 100) & 0xFFFFu;
  return exposed;
}
// Octree.java:651-670: only the 27 probes {c-1, c+s, c+s+1}^3
__device__ bool big_node_exposed(const Pyr &P, int ox, int oy, int oz, int cx, int cy, int cz, int s) {
  const int xs[3] = {cx - 1, cx + s, cx + s + 1}, ys[3] = {cy - 1, cy + s, cy + s + 1}, zs[3] = {cz - 1, cz + s, cz + s + 1};
  for (int a = 0; a < 3; a++) {
    if (zs[a] < 0 || zs[a] >= P.chunk) continue;
    for (int b = 0; b < 3; b++) {
      if (ys[b] < 0 || ys[b] >= P.chunk) continue;
      for (int c = 0; c < 3; c++) {
        if (xs[c] < 0 || xs[c] >= P.chunk) continue;
        if (voxel(P, ox + xs[c], oy + ys[b], oz + zs[a]) == 0u) return true;
      }
    }
  }
  return false;
}

__global__ void k_scale_heights(const uint16_t *__restrict__ height, uint16_t *__restrict__ h0, size_t nn, uint32_t quarter) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nn) h0[i] = (uint16_t)(((uint32_t)height[i] * quarter) >> 16);  // heightSample (svo_build.cpp)
}
__global__ void k_pyramid(const uint16_t *__restrict__ plo, const uint16_t *__restrict__ phi, const uint8_t *__restrict__ qlo,
                          const uint8_t *__restrict__ qhi, uint16_t *__restrict__ hlo, uint16_t *__restrict__ hhi, uint8_t *__restrict__ mlo,
                          uint8_t *__restrict__ mhi, uint32_t m) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)m * m) return;
  const uint32_t x = (uint32_t)(i % m), z = (uint32_t)(i / m);
  const size_t pm = (size_t)m * 2u, a = (size_t)(2u * z) * pm + 2u * x, b = a + 1, c = a + pm, d = c + 1;
  hlo[(size_t)z * m + x] = min(min(plo[a], plo[b]), min(plo[c], plo[d]));
  hhi[(size_t)z * m + x] = max(max(phi[a], phi[b]), max(phi[c], phi[d]));
  mlo[(size_t)z * m + x] = min(min(qlo[a], qlo[b]), min(qlo[c], qlo[d]));
  mhi[(size_t)z * m + x] = max(max(qhi[a], qhi[b]), max(qhi[c], qhi[d]));
}

// ---- sweep A: classify the eight children of every node of one level ------------------------------------------
// node size = 2 * cs; last = the children are single voxels (curLOD + 1 == maxLOD)
__global__ void __launch_bounds__(256) k_classify(Pyr P, Level lv, const int3 *__restrict__ chunk_origin, int cs, int Lc, bool last) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 3, n = t & 7u;
  const bool live = i < lv.n;
  uint32_t type = T_NONSURF, value = 0u;
  if (live) {
    const uint32_t k0 = lv.key0[i], k1 = lv.key1[i];
    const int3 o = chunk_origin[(k1 >> 16) >> 3];
    const int cx = (int)(k0 & 0xFFFFu) + (int)(n & 1u) * cs, cy = (int)(k0 >> 16) + (int)((n >> 1) & 1u) * cs,
              cz = (int)(k1 & 0xFFFFu) + (int)((n >> 2) & 1u) * cs;
    bool leaf = true;
    if (!last) classify(P, o.x + cx, o.y + cy, o.z + cz, cs, Lc, leaf, value);
    else value = voxel(P, o.x + cx, o.y + cy, o.z + cz);
    // Octree.java:556-600
    if (leaf && value != 0u) {
      if (cs == 1) {
        uint32_t packed;
        type = surface_normal(P, o.x, o.y, o.z, cx, cy, cz, packed) ? T_SURFACE : T_NONSURF;
      } else {
        type = big_node_exposed(P, o.x, o.y, o.z, cx, cy, cz, cs) ? T_INTERIOR : T_SUBDIV;
      }
    } else if (leaf) {
      type = cs == 1 ? T_NONSURF : T_SUBDIV;
    } else {
      type = T_INTERIOR;
    }
  }
  // assemble the node's words from its eight lanes (lanes 8g .. 8g+7 of the warp)
  const uint32_t size = type == T_SURFACE ? 3u : (type == T_NONSURF ? 1u : 7u);
  const bool rec = live && !last && type == T_INTERIOR && value != 0u;  // constructInnerOctree recurses (Octree.java:603-607)
  uint32_t codes = type << (2u * n), block = size, recm = rec ? (1u << n) : 0u;
  uint32_t vlo = n < 4u ? value << (8u * n) : 0u, vhi = n >= 4u ? value << (8u * (n - 4u)) : 0u;
#pragma unroll
  for (int d = 1; d < 8; d <<= 1) {
    codes |= __shfl_xor_sync(0xffffffffu, codes, d);
    block += __shfl_xor_sync(0xffffffffu, block, d);
    recm |= __shfl_xor_sync(0xffffffffu, recm, d);
    vlo |= __shfl_xor_sync(0xffffffffu, vlo, d);
    vhi |= __shfl_xor_sync(0xffffffffu, vhi, d);
  }
  if (live && n == 0u) {
    lv.packed[i] = codes | (block << 16) | (recm << 24);
    lv.vals[i] = make_uint2(vlo, vhi);
    lv.first[i] = (uint32_t)__popc(recm);  // count; scanned in place afterwards
  }
}

// the next level's node list: recursing children in child order, at the slots the scan assigned
__global__ void __launch_bounds__(256) k_expand(Level lv, Level nx, int cs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lv.n) return;
  const uint32_t recm = lv.packed[i] >> 24;
  if (recm == 0u) return;
  const uint32_t k0 = lv.key0[i], k1 = lv.key1[i];
  uint32_t slot = lv.first[i];
  for (uint32_t n = 0; n < 8u; n++)
    if ((recm >> n) & 1u) {
      const uint32_t cx = (k0 & 0xFFFFu) + (n & 1u) * (uint32_t)cs, cy = (k0 >> 16) + ((n >> 1) & 1u) * (uint32_t)cs,
                     cz = (k1 & 0xFFFFu) + ((n >> 2) & 1u) * (uint32_t)cs;
      nx.key0[slot] = cx | (cy << 16);
      nx.key1[slot] = cz | (k1 & 0xFFFF0000u);
      slot++;
    }
}

// ---- sweep B: subtree bytes, deepest level first ------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sizes(Level lv, const uint32_t *__restrict__ childS) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lv.n) return;
  const uint32_t p = lv.packed[i];
  uint32_t s = (p >> 16) & 0xFFu;
  const uint32_t cnt = (uint32_t)__popc(p >> 24), f = lv.first[i];
  for (uint32_t r = 0; r < cnt; r++) s += childS[f + r];
  lv.S[i] = s;
}

// ---- sweep C: offsets (top-down), then the bytes ------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_offsets(Level lv, Level nx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lv.n) return;
  const uint32_t p = lv.packed[i], recm = p >> 24;
  if (recm == 0u) return;
  const uint32_t off = lv.off[i];
  uint32_t childpos = off, running = off + ((p >> 16) & 0xFFu), c = lv.first[i];
  for (uint32_t n = 0; n < 8u; n++) {
    const uint32_t type = (p >> (2u * n)) & 3u;
    if ((recm >> n) & 1u) {
      nx.recpos[c] = childpos;
      nx.off[c] = running;
      running += nx.S[c];
      c++;
    }
    childpos += type == T_SURFACE ? 3u : (type == T_NONSURF ? 1u : 7u);
  }
}

__global__ void __launch_bounds__(256) k_emit(Pyr P, Level lv, const int3 *__restrict__ chunk_origin, int cs, uint8_t *__restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 3, n = t & 7u;
  if (i >= lv.n) return;
  const uint32_t p = lv.packed[i];
  const uint2 v = lv.vals[i];
  const uint32_t off = lv.off[i];
  // offset of child n inside the block: 7 / 3 / 1 bytes per earlier sibling
  uint32_t rel = 0;
  for (uint32_t m = 0; m < n; m++) {
    const uint32_t ty = (p >> (2u * m)) & 3u;
    rel += ty == T_SURFACE ? 3u : (ty == T_NONSURF ? 1u : 7u);
  }
  const uint32_t type = (p >> (2u * n)) & 3u;
  const uint32_t value = ((n < 4u ? v.x >> (8u * n) : v.y >> (8u * (n - 4u)))) & 0xFFu;
  uint8_t *rec = out + (size_t)off + rel;
  rec[0] = (uint8_t)value;  // the other bytes of 7-byte records stay zero unless the child writes them below (it is a node of the next level)
  if (type == T_SURFACE) {
    const uint32_t k0 = lv.key0[i], k1 = lv.key1[i];
    const int3 o = chunk_origin[(k1 >> 16) >> 3];
    const int cx = (int)(k0 & 0xFFFFu) + (int)(n & 1u) * cs, cy = (int)(k0 >> 16) + (int)((n >> 1) & 1u) * cs,
              cz = (int)(k1 & 0xFFFFu) + (int)((n >> 2) & 1u) * cs;
    uint32_t packed = 0;
    surface_normal(P, o.x, o.y, o.z, cx, cy, cz, packed);
    rec[1] = (uint8_t)(packed & 0xFFu);  // little-endian (Octree.java:141-150)
    rec[2] = (uint8_t)(packed >> 8);
  }
  if (n == 0u) {  // this node's own record: child pointer relative to the record (big-endian) and the type mask
    const uint32_t rp = lv.recpos[i], cp = off - rp, mask = p & 0xFFFFu;
    uint8_t *own = out + (size_t)rp;
    own[1] = (uint8_t)(cp >> 24); own[2] = (uint8_t)(cp >> 16); own[3] = (uint8_t)(cp >> 8); own[4] = (uint8_t)cp;
    own[5] = (uint8_t)(mask >> 8); own[6] = (uint8_t)mask;
  }
}

__global__ void k_put_values(uint8_t *__restrict__ out, const uint32_t *__restrict__ pos, uint32_t n, uint8_t value) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[pos[i]] = value;
}

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return l;
}
const int kOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};  // Octree.java:42-51

struct HostTree {  // the nodes above the chunks (Octree.java:234-244, fillEmptyChildren :481-502)
  std::vector<uint8_t> buf;
  size_t off = 0;
  size_t put7(uint8_t v) {
    if (buf.size() < off + 7) buf.resize(off + 7, 0);
    buf[off] = v;
    const size_t p = off;
    off += 7;
    return p;
  }
  void set_cp(size_t parent, uint64_t child) {
    const uint32_t rel = (uint32_t)((int64_t)child - (int64_t)parent);
    buf[parent + 1] = (uint8_t)(rel >> 24); buf[parent + 2] = (uint8_t)(rel >> 16);
    buf[parent + 3] = (uint8_t)(rel >> 8);  buf[parent + 4] = (uint8_t)rel;
  }
};
struct ChunkRef { int origin[3]; size_t pointer; };
void fill_empty(HostTree &t, size_t parent, int levels, const int p[3], int chunk, std::vector<ChunkRef> &chunks) {
  if (levels == 0) {
    ChunkRef c;
    c.origin[0] = p[0]; c.origin[1] = p[1]; c.origin[2] = p[2];
    c.pointer = parent;
    chunks.push_back(c);
    return;
  }
  const int cs = chunk << (levels - 1);
  size_t children[8];
  int cp[8][3];
  for (int n = 0; n < 8; n++)
    for (int a = 0; a < 3; a++) cp[n][a] = p[a] + kOff[n][a] * cs;
  for (int i = 0; i < 8; i++) children[i] = t.put7(1);
  for (int i = 0; i < 8; i++) fill_empty(t, children[i], levels - 1, cp[i], chunk, chunks);
  t.set_cp(parent, children[0]);
}

#define GB_CUDA(call)                    \
  do {                                   \
    cudaError_t e_ = (call);             \
    if (e_ != cudaSuccess) return e_;    \
  } while (0)

cudaError_t alloc_level(Pool &pool, Level &lv, uint32_t n) {
  lv.n = n;
  GB_CUDA(pool.get(&lv.key0, n)); GB_CUDA(pool.get(&lv.key1, n)); GB_CUDA(pool.get(&lv.packed, n)); GB_CUDA(pool.get(&lv.vals, n));
  GB_CUDA(pool.get(&lv.first, (size_t)n + 1)); GB_CUDA(pool.get(&lv.S, n)); GB_CUDA(pool.get(&lv.off, n)); GB_CUDA(pool.get(&lv.recpos, n));
  return cudaSuccess;
}

}  // namespace

// height / mat: n x n on the HOST (row = z).  *stream receives a cudaMalloc'ed buffer of *capacity bytes holding the node
// stream in [0, *nbytes) (the caller owns it).  unsupported = true (and nothing allocated) for shapes this path does not
// take (more than 65535 sub-octrees, chunk > 32768): the caller falls back to the host builder.
cudaError_t gpu_build_terrain(const uint16_t *height, const uint8_t *mat, int n, int chunk, uint8_t **stream, uint64_t *nbytes,
                              uint64_t *capacity, bool *unsupported, cudaStream_t st, uint64_t *launches) {
  *unsupported = false;
  *stream = nullptr;
  if (chunk > n) chunk = n;
  const int lgn = ilog2(n), levels = ilog2(n / chunk);
  HostTree t;
  std::vector<ChunkRef> chunks;
  const int root_pos[3] = {0, 0, 0};
  t.put7(1);
  fill_empty(t, 0, levels, root_pos, chunk, chunks);
  const size_t nunits = chunks.size() * 8;
  if (nunits > 65535 || chunk > 32768 || chunk < 4) { *unsupported = true; return cudaSuccess; }
  uint64_t nl = 0;

  Pool pool;
  // ---- maps and pyramids ----
  const size_t nn = (size_t)n * (size_t)n;
  uint16_t *d_height = nullptr;
  Pyr P;
  P.n = n;
  P.chunk = chunk;
  uint16_t *hlo[16], *hhi[16];
  uint8_t *mlo[16], *mhi[16];
  GB_CUDA(pool.get(&d_height, nn));
  GB_CUDA(pool.get(&hlo[0], nn));
  GB_CUDA(pool.get(&mlo[0], nn));
  hhi[0] = hlo[0];
  mhi[0] = mlo[0];
  GB_CUDA(dev_copy(d_height, height, nn * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
  GB_CUDA(dev_copy(mlo[0], mat, nn, cudaMemcpyHostToDevice, st));
  SVO_LAUNCH((unsigned)((nn + 255) / 256), 256, st, k_scale_heights)(d_height, hlo[0], nn, (uint32_t)(n / 4));
  nl++;
  for (int L = 1; L <= lgn; L++) {
    const size_t m = (size_t)n >> L;
    GB_CUDA(pool.get(&hlo[L], m * m)); GB_CUDA(pool.get(&hhi[L], m * m)); GB_CUDA(pool.get(&mlo[L], m * m)); GB_CUDA(pool.get(&mhi[L], m * m));
    SVO_LAUNCH((unsigned)((m * m + 255) / 256), 256, st, k_pyramid)(hlo[L - 1], hhi[L - 1], mlo[L - 1], mhi[L - 1], hlo[L], hhi[L], mlo[L], mhi[L], (uint32_t)m);
    nl++;
  }
  for (int L = 0; L < 16; L++) {
    P.hmin[L] = L <= lgn ? hlo[L] : nullptr; P.hmax[L] = L <= lgn ? hhi[L] : nullptr;
    P.mmin[L] = L <= lgn ? mlo[L] : nullptr; P.mmax[L] = L <= lgn ? mhi[L] : nullptr;
  }
  std::vector<int3> h_origin(chunks.size());
  for (size_t c = 0; c < chunks.size(); c++) h_origin[c] = make_int3(chunks[c].origin[0], chunks[c].origin[1], chunks[c].origin[2]);
  int3 *d_origin = nullptr;
  GB_CUDA(pool.get(&d_origin, chunks.size()));
  GB_CUDA(dev_copy(d_origin, h_origin.data(), chunks.size() * sizeof(int3), cudaMemcpyHostToDevice, st));

  // ---- sweep A ----
  // unit roots: constructInnerOctree(size = chunk / 2, curLOD 0, maxLOD = log2(chunk / 2)) at kOff[i] * chunk / 2 (OctreeThread.java:19-23)
  const int unit_size = chunk / 2, sub_lod = ilog2(unit_size);
  std::vector<Level> lv;
  if (sub_lod >= 1) {
    Level l0;
    GB_CUDA(alloc_level(pool, l0, (uint32_t)nunits));
    std::vector<uint32_t> k0(nunits), k1(nunits);
    for (size_t u = 0; u < nunits; u++) {
      const int i = (int)(u % 8);
      k0[u] = (uint32_t)(kOff[i][0] * unit_size) | ((uint32_t)(kOff[i][1] * unit_size) << 16);
      k1[u] = (uint32_t)(kOff[i][2] * unit_size) | ((uint32_t)u << 16);
    }
    GB_CUDA(dev_copy(l0.key0, k0.data(), nunits * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(dev_copy(l0.key1, k1.data(), nunits * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(dev_sync(st));  // k0 / k1 go out of scope
    lv.push_back(l0);
  }
  void *d_scan_tmp = nullptr;
  size_t scan_bytes = 0;
  for (int k = 0; k < sub_lod; k++) {
    Level &cur = lv[k];
    const int cs = unit_size >> (k + 1);  // child size at this level
    const bool last = (k + 1 == sub_lod);
    if (cur.n) {
      SVO_LAUNCH((unsigned)(((size_t)cur.n * 8 + 255) / 256), 256, st, k_classify)(P, cur, d_origin, cs, ilog2(cs), last);
      nl++;
    }
    if (last) break;
    // counts -> slots (exclusive scan over n + 1 entries: the last one becomes the total)
    GB_CUDA(dev_memset(cur.first + cur.n, 0, sizeof(uint32_t), st));
    size_t need = 0;
    GB_CUDA(exclusive_scan(nullptr, need, cur.first, (int)cur.n + 1, st));  // size query
    if (need > scan_bytes) {
      GB_CUDA(pool.get((uint8_t **)&d_scan_tmp, need));
      scan_bytes = need;
    }
    GB_CUDA(exclusive_scan(d_scan_tmp, need, cur.first, (int)cur.n + 1, st));
    nl++;
    uint32_t total = 0;
    GB_CUDA(dev_copy(&total, cur.first + cur.n, sizeof total, cudaMemcpyDeviceToHost, st));
    GB_CUDA(dev_sync(st));
    Level nx;
    GB_CUDA(alloc_level(pool, nx, total));
    if (cur.n && total) {
      SVO_LAUNCH((cur.n + 255) / 256, 256, st, k_expand)(cur, nx, cs);
      nl++;
    }
    lv.push_back(nx);
  }
  // ---- sweep B ----
  for (int k = (int)lv.size() - 1; k >= 0; k--)
    if (lv[k].n) {
      SVO_LAUNCH((lv[k].n + 255) / 256, 256, st, k_sizes)(lv[k], k + 1 < (int)lv.size() ? lv[k + 1].S : nullptr);
      nl++;
    }
  // ---- layout of everything above the sub-octrees (Octree.java:317-337), from the unit sizes ----
  std::vector<uint32_t> unitS(nunits, 0u);
  if (!lv.empty()) GB_CUDA(dev_copy(unitS.data(), lv[0].S, nunits * 4, cudaMemcpyDeviceToHost, st));
  GB_CUDA(dev_sync(st));
  uint64_t off = t.off;
  std::vector<uint32_t> u_recpos(nunits), u_off(nunits);
  for (size_t ci = 0; ci < chunks.size(); ci++) {
    const uint64_t children0 = off;
    off += 56;                              // the chunk's eight interior children (value 1)
    t.set_cp(chunks[ci].pointer, children0);
    for (int i = 0; i < 8; i++) {
      const size_t u = ci * 8 + (size_t)i;
      u_recpos[u] = (uint32_t)(children0 + 7u * (uint64_t)i);  // the unit's dummy head lands in the chunk's child record
      u_off[u] = (uint32_t)off;             // its child block starts here (the head itself is not copied)
      off += 7u + (uint64_t)unitS[u];       // memOffset += childOffset: 7 bytes more than were copied (Octree.java:336)
    }
  }
  const uint64_t total_bytes = off;
  if (total_bytes >= (1ull << 32)) { *unsupported = true; return cudaSuccess; }
  const uint64_t cap = total_bytes + total_bytes / 16 + 4096;
  uint8_t *out = nullptr;
  GB_CUDA(dev_alloc((void **)&out, cap));
  cudaError_t e = cudaSuccess;
  do {
    if ((e = dev_memset(out, 0, cap, st)) != cudaSuccess) break;
    if ((e = dev_copy(out, t.buf.data(), t.off, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if (!lv.empty()) {
      if ((e = dev_copy(lv[0].recpos, u_recpos.data(), nunits * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      if ((e = dev_copy(lv[0].off, u_off.data(), nunits * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      SVO_LAUNCH((unsigned)((nunits + 255) / 256), 256, st, k_put_values)(out, lv[0].recpos, (uint32_t)nunits, 1);  // chunk children: value 1
      nl++;
    } else {  // chunk of 2 voxels: the units have no levels; still the chunk children exist
      uint32_t *d_pos = nullptr;
      if ((e = pool.get(&d_pos, nunits)) != cudaSuccess) break;
      if ((e = dev_copy(d_pos, u_recpos.data(), nunits * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      SVO_LAUNCH((unsigned)((nunits + 255) / 256), 256, st, k_put_values)(out, d_pos, (uint32_t)nunits, 1);
      nl++;
    }
    // ---- sweep C ----
    for (size_t k = 0; k < lv.size(); k++) {
      if (!lv[k].n) continue;
      if (k + 1 < lv.size() && lv[k + 1].n) {
        SVO_LAUNCH((lv[k].n + 255) / 256, 256, st, k_offsets)(lv[k], lv[k + 1]);
        nl++;
      }
      const int cs = unit_size >> (k + 1);
      SVO_LAUNCH((unsigned)(((size_t)lv[k].n * 8 + 255) / 256), 256, st, k_emit)(P, lv[k], d_origin, cs, out);
      nl++;
    }
    if ((e = dev_last_error()) != cudaSuccess) break;
    e = dev_sync(st);  // host vectors above are still being read by the copies
  } while (0);
  if (e != cudaSuccess) {
    dev_free(out);
    return e;
  }
  *stream = out;
  *nbytes = total_bytes;
  *capacity = cap;
  if (launches) *launches = nl;
  return cudaSuccess;
}

}  // namespace svo
