// svo_build.cpp -- heightmap world generator producing the engine's node
// stream byte-for-byte (svo_build_terrain in include/svo_b200.h).
//
// What: the stream Octree.constructCompleteOctree writes for a heightmap world
// (reference src/engine/Octree.java:192-353: root + fillEmptyChildren levels,
// then per 1024^3 chunk eight 512^3 sub-octrees built by OctreeThread.java and
// spliced in), with the voxel rule of src/shaders/chunkgen-heightmap.comp:13-31.
//
// How (new): the reference materialises every chunk as a dense 1 GiB voxel
// image on the GPU, reads it back and scans O(size^3) voxels per node.  A
// heightmap world is fully described by its column heights, so this builder
// never materialises voxels: node homogeneity comes from min/max mip pyramids
// of the height and material maps (O(1) for almost every node, O(size^2)
// columns for the few that straddle the 5-voxel material band), the scan-order
// dependent `value` of mixed nodes (Octree.java:527-547) is recovered from the
// first column reaching the node's floor, and all chunk x 8 sub-octrees are
// built in parallel (their child pointers are parent-relative, so they splice
// by plain copy exactly like Octree.java:322-337, including its 7-byte gap).
// An 8192^3 world (512 GiB of voxels upstream) builds in seconds per core.
#include "../../include/svo_b200.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct World {
  int n = 0, lgn = 0;
  std::vector<std::vector<uint16_t>> hmin, hmax;  // [level][ (n>>level)^2 ], level 0 = column heights
  std::vector<std::vector<uint8_t>> mmin, mmax;   // material ids
  inline int h(int wx, int wz) const { return hmin[0][(size_t)wz * (size_t)n + (size_t)wx]; }
  inline int m(int wx, int wz) const { return mmin[0][(size_t)wz * (size_t)n + (size_t)wx]; }
  // chunkgen-heightmap.comp:22-28
  inline uint8_t voxel(int wx, int wy, int wz) const {
    const int hs = h(wx, wz);
    if (wy > hs) return 0;
    return (hs - wy <= 4) ? (uint8_t)m(wx, wz) : (uint8_t)1;
  }
};

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return l;
}

const int kOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};  // Octree.java:42-51

enum { T_INTERIOR = 0, T_SURFACE = 1, T_SUBDIV = 2, T_NONSURF = 3 };

struct Unit {  // one OctreeThread's private Octree (OctreeThread.java:19-23)
  const World *w;
  int chunk;      // CHUNK_SIZE
  int ox, oy, oz; // world origin of the chunk
  std::vector<uint8_t> buf;

  size_t put(int nbytes, uint8_t first) {
    size_t p = buf.size();
    buf.resize(p + (size_t)nbytes, 0);
    buf[p] = first;
    return p;
  }
  void set_child_pointer(size_t parent, size_t child) {  // Octree.java:162-164
    uint32_t rel = (uint32_t)((int64_t)child - (int64_t)parent);
    buf[parent + 1] = (uint8_t)(rel >> 24); buf[parent + 2] = (uint8_t)(rel >> 16);
    buf[parent + 3] = (uint8_t)(rel >> 8);  buf[parent + 4] = (uint8_t)rel;
  }
  void set_leaf_mask(size_t parent, uint16_t mask) {  // Octree.java:170-172
    buf[parent + 5] = (uint8_t)(mask >> 8); buf[parent + 6] = (uint8_t)mask;
  }
  inline uint8_t vox(int x, int y, int z) const { return w->voxel(ox + x, oy + y, oz + z); }

  // Octree.java:620-649
  bool surface_normal(int cx, int cy, int cz, int16_t &packed) const {
    bool exposed = false;
    int nx = 0, ny = 0, nz = 0;
    for (int i = cx - 1; i <= cx + 1; i++) {
      if (i < 0 || i >= chunk) continue;
      for (int k = cz - 1; k <= cz + 1; k++) {
        if (k < 0 || k >= chunk) continue;
        const int hs = w->h(ox + i, oz + k);
        for (int j = cy - 1; j <= cy + 1; j++) {
          if (j < 0 || j >= chunk) continue;
          if (oy + j > hs) { exposed = true; nx += i - cx; ny += j - cy; nz += k - cz; }  // voxel == 0 iff above the column
        }
      }
    }
    nx = nx / 2 + 5; ny = ny / 2 + 5; nz = nz / 2 + 5;
    packed = (int16_t)(nx + ny * 10 + nz * 100);
    return exposed;
  }
  // Octree.java:651-670: only the 27 probes {c-1, c+s, c+s+1}^3
  bool big_node_exposed(int cx, int cy, int cz, int s) const {
    const int xs[3] = {cx - 1, cx + s, cx + s + 1}, ys[3] = {cy - 1, cy + s, cy + s + 1}, zs[3] = {cz - 1, cz + s, cz + s + 1};
    for (int a = 0; a < 3; a++) {
      if (zs[a] < 0 || zs[a] >= chunk) continue;
      for (int b = 0; b < 3; b++) {
        if (ys[b] < 0 || ys[b] >= chunk) continue;
        for (int c = 0; c < 3; c++) {
          if (xs[c] < 0 || xs[c] >= chunk) continue;
          if (vox(xs[c], ys[b], zs[a]) == 0) return true;
        }
      }
    }
    return false;
  }

  // The scan of Octree.java:527-555 for the cube (cx,cy,cz,s): is it homogeneous, and its `value`.
  void classify(int cx, int cy, int cz, int s, bool &leaf, uint8_t &value) const {
    const int L = ilog2(s);
    const int wx0 = ox + cx, wy0 = oy + cy, wz0 = oz + cz;
    const size_t pi = (size_t)(wz0 >> L) * (size_t)(w->n >> L) + (size_t)(wx0 >> L);
    const int hmn = w->hmin[L][pi], hmx = w->hmax[L][pi];
    const int y1 = wy0 + s - 1;
    const uint8_t first = w->voxel(wx0, wy0, wz0);
    if (wy0 > hmx) { leaf = true; value = 0; return; }  // all air
    if (y1 > hmn) {                                      // air above the lowest column, rock in the highest
      leaf = false;
      if (first != 0) { value = first; return; }
      // first == 0: `first` becomes the first non-zero sample in scan order z, y, x.  A slab z has a solid
      // voxel iff some column reaches wy0, and then the first one met is on the row y = wy0.
      for (int z = wz0; z < wz0 + s; z++)
        for (int x = wx0; x < wx0 + s; x++) {
          const int hs = w->h(x, z);
          if (hs >= wy0) { value = (hs - wy0 <= 4) ? (uint8_t)w->m(x, z) : (uint8_t)1; return; }
        }
      value = 0;  // unreachable (hmx >= wy0)
      return;
    }
    // every voxel is solid
    value = first;
    if (y1 < hmn - 4) { leaf = true; return; }  // below every material band: all 1
    if (w->mmin[L][pi] == 1 && w->mmax[L][pi] == 1) { leaf = true; return; }
    for (int z = wz0; z < wz0 + s; z++)
      for (int x = wx0; x < wx0 + s; x++) {
        const int hs = w->h(x, z);
        const bool band = hs - 4 <= y1;    // some y in range with hs - y <= 4 -> material
        const bool deep = wy0 <= hs - 5;   // some y in range with hs - y >= 5 -> 1
        if ((band && (uint8_t)w->m(x, z) != first) || (deep && first != 1)) { leaf = false; return; }
      }
    leaf = true;
  }

  // Octree.java:511-608
  void construct(int size, int curLOD, int maxLOD, int px, int py, int pz, size_t parent) {
    const int cs = size / 2;
    if (cs == 0 || curLOD == maxLOD) return;
    size_t children[8];
    int types[8], cp[8][3];
    uint16_t leaf_mask = 0;
    for (int n = 0; n < 8; n++) {
      cp[n][0] = px + kOff[n][0] * cs; cp[n][1] = py + kOff[n][1] * cs; cp[n][2] = pz + kOff[n][2] * cs;
      bool leaf = true;
      uint8_t value;
      if (curLOD + 1 != maxLOD) classify(cp[n][0], cp[n][1], cp[n][2], cs, leaf, value);
      else value = vox(cp[n][0], cp[n][1], cp[n][2]);
      int type;
      if (leaf && value != 0) {
        if (cs == 1) {
          int16_t normal;
          if (surface_normal(cp[n][0], cp[n][1], cp[n][2], normal)) {
            children[n] = put(3, value);
            buf[children[n] + 1] = (uint8_t)(normal & 0xff);
            buf[children[n] + 2] = (uint8_t)((normal >> 8) & 0xff);
            type = T_SURFACE;
          } else { children[n] = put(1, value); type = T_NONSURF; }
        } else if (big_node_exposed(cp[n][0], cp[n][1], cp[n][2], cs)) { children[n] = put(7, value); type = T_INTERIOR; }
        else { children[n] = put(7, value); type = T_SUBDIV; }
      } else if (leaf) {
        if (cs == 1) { children[n] = put(1, value); type = T_NONSURF; }
        else { children[n] = put(7, value); type = T_SUBDIV; }
      } else { children[n] = put(7, value); type = T_INTERIOR; }
      types[n] = type;
      leaf_mask |= (uint16_t)(type << (n << 1));
    }
    set_child_pointer(parent, children[0]);
    set_leaf_mask(parent, leaf_mask);
    for (int n = 0; n < 8; n++)
      if (buf[children[n]] != 0 && types[n] == T_INTERIOR) construct(cs, curLOD + 1, maxLOD, cp[n][0], cp[n][1], cp[n][2], children[n]);
  }
};

struct ChunkRef { int origin[3]; size_t pointer; };

struct MainTree {
  std::vector<uint8_t> buf;
  size_t off = 0;  // memOffset
  size_t put7(uint8_t v) {
    if (buf.size() < off + 7) buf.resize(off + 7, 0);
    buf[off] = v;
    size_t p = off;
    off += 7;
    return p;
  }
  void set_child_pointer(size_t parent, size_t child) {
    uint32_t rel = (uint32_t)((int64_t)child - (int64_t)parent);
    buf[parent + 1] = (uint8_t)(rel >> 24); buf[parent + 2] = (uint8_t)(rel >> 16);
    buf[parent + 3] = (uint8_t)(rel >> 8);  buf[parent + 4] = (uint8_t)rel;
  }
};

// Octree.java:481-502
void fill_empty(MainTree &t, size_t parent, int levels, const int p[3], int chunk, std::vector<ChunkRef> &chunks) {
  if (levels == 0) {
    ChunkRef c;
    memcpy(c.origin, p, sizeof c.origin);
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
  t.set_child_pointer(parent, children[0]);
}

}  // namespace

extern "C" int svo_build_terrain(const uint16_t *height, const uint8_t *mat, int n, int chunk, uint8_t *out, uint64_t cap,
                                 uint64_t *out_bytes, int nthreads) {
  if (!height || !mat || !out_bytes || n < 2 || (n & (n - 1)) || chunk < 2 || (chunk & (chunk - 1))) return SVO_ERR_INVALID;
  if (chunk > n) chunk = n;
  if (nthreads < 1) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;

  World w;
  w.n = n;
  w.lgn = ilog2(n);
  const int levels_pyr = w.lgn + 1;
  w.hmin.resize(levels_pyr); w.hmax.resize(levels_pyr); w.mmin.resize(levels_pyr); w.mmax.resize(levels_pyr);
  const size_t nn = (size_t)n * (size_t)n;
  w.hmin[0].resize(nn);
  w.mmin[0].assign(mat, mat + nn);
  for (size_t i = 0; i < nn; i++) w.hmin[0][i] = (uint16_t)(((uint32_t)height[i] * (uint32_t)(n / 4)) >> 16);  // heightSample
  for (int L = 1; L < levels_pyr; L++) {
    const size_t m = (size_t)n >> L, pm = (size_t)n >> (L - 1);
    w.hmin[L].resize(m * m); w.hmax[L].resize(m * m); w.mmin[L].resize(m * m); w.mmax[L].resize(m * m);
    const std::vector<uint16_t> &plo = w.hmin[L - 1], &phi = (L == 1 ? w.hmin[0] : w.hmax[L - 1]);
    const std::vector<uint8_t> &qlo = w.mmin[L - 1], &qhi = (L == 1 ? w.mmin[0] : w.mmax[L - 1]);
    for (size_t z = 0; z < m; z++)
      for (size_t x = 0; x < m; x++) {
        const size_t a = (2 * z) * pm + 2 * x, b = a + 1, c = a + pm, d = c + 1;
        uint16_t lo = plo[a], hi = phi[a];
        lo = std::min(std::min(lo, plo[b]), std::min(plo[c], plo[d]));
        hi = std::max(std::max(hi, phi[b]), std::max(phi[c], phi[d]));
        w.hmin[L][z * m + x] = lo; w.hmax[L][z * m + x] = hi;
        uint8_t ml = qlo[a], mh = qhi[a];
        ml = std::min(std::min(ml, qlo[b]), std::min(qlo[c], qlo[d]));
        mh = std::max(std::max(mh, qhi[b]), std::max(qhi[c], qhi[d]));
        w.mmin[L][z * m + x] = ml; w.mmax[L][z * m + x] = mh;
      }
  }
  w.hmax[0].clear();  // level 0: hmin doubles as hmax (classify never asks level 0: cs >= 2 there)
  w.hmax[0] = w.hmin[0];
  w.mmax[0] = w.mmin[0];

  MainTree t;
  const int levels = ilog2(n / chunk);
  std::vector<ChunkRef> chunks;
  const int root_pos[3] = {0, 0, 0};
  t.put7(1);                                          // Octree.java:234
  fill_empty(t, 0, levels, root_pos, chunk, chunks);  // :244

  // all chunk x 8 sub-octrees in parallel (upstream: 8 threads per chunk, chunks serial)
  const size_t nunits = chunks.size() * 8;
  std::vector<std::vector<uint8_t>> units(nunits);
  std::atomic<size_t> next(0);
  const int cs = chunk / 2, sub_lod = ilog2(cs);
  auto worker = [&]() {
    for (;;) {
      const size_t u = next.fetch_add(1);
      if (u >= nunits) return;
      const ChunkRef &c = chunks[u / 8];
      const int i = (int)(u % 8);
      Unit unit;
      unit.w = &w; unit.chunk = chunk;
      unit.ox = c.origin[0]; unit.oy = c.origin[1]; unit.oz = c.origin[2];
      unit.buf.reserve(1 << 16);
      unit.put(7, 1);  // createDummyHead
      unit.construct(cs, 0, sub_lod, kOff[i][0] * cs, kOff[i][1] * cs, kOff[i][2] * cs, 0);
      units[u].swap(unit.buf);
    }
  };
  std::vector<std::thread> pool;
  for (int k = 1; k < nthreads; k++) pool.emplace_back(worker);
  worker();
  for (auto &th : pool) th.join();

  // splice (Octree.java:317-337)
  uint64_t total = t.off;
  for (size_t u = 0; u < nunits; u++) total += (u % 8 == 0 ? 56 : 0) + units[u].size();
  *out_bytes = total;
  if (!out) return SVO_OK;
  if (cap < total) return SVO_ERR_OOM;
  memset(out, 0, total);
  memcpy(out, t.buf.data(), t.off);
  uint64_t off = t.off;
  auto put7 = [&](uint8_t v) { out[off] = v; uint64_t p = off; off += 7; return p; };
  auto set_cp = [&](uint64_t parent, uint64_t child) {
    uint32_t rel = (uint32_t)((int64_t)child - (int64_t)parent);
    out[parent + 1] = (uint8_t)(rel >> 24); out[parent + 2] = (uint8_t)(rel >> 16);
    out[parent + 3] = (uint8_t)(rel >> 8);  out[parent + 4] = (uint8_t)rel;
  };
  for (size_t ci = 0; ci < chunks.size(); ci++) {
    uint64_t children[8];
    for (int i = 0; i < 8; i++) children[i] = put7(1);
    set_cp(chunks[ci].pointer, children[0]);
    for (int i = 0; i < 8; i++) {
      const std::vector<uint8_t> &sub = units[ci * 8 + i];
      set_cp(children[i], off);
      out[children[i] + 5] = sub[5];  // head leaf mask
      out[children[i] + 6] = sub[6];
      memcpy(out + off, sub.data() + 7, sub.size() - 7);
      off += sub.size();  // memOffset += childOffset: 7 bytes more than were copied (Octree.java:336)
    }
  }
  return SVO_OK;
}
