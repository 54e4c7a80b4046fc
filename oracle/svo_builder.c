/*
 * svo_builder.c -- TEST INFRASTRUCTURE.  Brute-force CPU restatement of the
 * reference octree builder, /root/reference/src/engine/Octree.java, used to
 * (a) make reference-format node streams for the traversal oracle and
 * (b) check the product's fast builder byte-for-byte.
 *
 * PINNED: no .svo level file ships with the reference (README.md:10) and no JVM
 * exists here, but the builder's Java text compiles as C++ after a mechanical
 * rewrite (oracle/build_ref_java.py -> oracle/_ref/libsvo_ref_java.so);
 * tests/test_builder_ref.py demands byte-equal streams from that library and
 * from this file (dense volumes, heightmap worlds with chunk splices).
 *
 * Follows: record writers Octree.java:119-176, fillEmptyChildren :481-502,
 * constructInnerOctree :511-608, genSurfaceNormal :620-649,
 * checkBigNodeExposed :651-670, chunk splice :292-343, OctreeThread.java:19-23,
 * voxel rule src/shaders/chunkgen-heightmap.comp:13-31.
 * Generalisation: CHUNK_SIZE (1024 upstream) is the `chunk` argument and the
 * world edge n = chunk << levels (upstream: chunk 1024, levels 3).
 */
#include "svo_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
  uint8_t *buf;
  uint64_t cap, off;
  int overflow;
  uint64_t counts[4]; /* surface, non-surface, subdividable, interior */
  const uint8_t *vox; /* dense chunk, index x | y<<lg | z<<2lg (Octree.java:110-112) */
  int lg, chunk;
} oct_t;

static const int child_offsets[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0},
                                        {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}}; /* :42-51 */

static inline uint8_t get_voxel(const oct_t *o, int x, int y, int z) {
  return o->vox[(size_t)x | ((size_t)y << o->lg) | ((size_t)z << (2 * o->lg))];
}

static uint64_t put_bytes(oct_t *o, int n, uint8_t first) {
  uint64_t p = o->off;
  if (p + (uint64_t)n > o->cap) { o->overflow = 1; return 0; }
  o->buf[p] = first;
  for (int i = 1; i < n; i++) o->buf[p + (uint64_t)i] = 0;
  o->off += (uint64_t)n;
  return p;
}
static uint64_t create_interior(oct_t *o, uint8_t v) { o->counts[3]++; return put_bytes(o, 7, v); }     /* :119-130 */
static uint64_t create_subdividable(oct_t *o, uint8_t v) { o->counts[2]++; return put_bytes(o, 7, v); } /* :132-144 */
static uint64_t create_surface(oct_t *o, uint8_t v, int16_t normal) {                                    /* :146-153 */
  o->counts[0]++;
  uint64_t p = put_bytes(o, 3, v);
  if (!o->overflow) { o->buf[p + 1] = (uint8_t)(normal & 0xff); o->buf[p + 2] = (uint8_t)((normal >> 8) & 0xff); }
  return p;
}
static uint64_t create_non_surface(oct_t *o, uint8_t v) { o->counts[1]++; return put_bytes(o, 1, v); }   /* :155-160 */

static void set_child_pointer(oct_t *o, uint64_t parent, uint64_t child) {                               /* :162-164 */
  if (o->overflow) return;
  uint32_t rel = (uint32_t)((int64_t)child - (int64_t)parent);
  o->buf[parent + 1] = (uint8_t)(rel >> 24); o->buf[parent + 2] = (uint8_t)(rel >> 16);
  o->buf[parent + 3] = (uint8_t)(rel >> 8);  o->buf[parent + 4] = (uint8_t)rel;
}
static void set_leaf_mask(oct_t *o, uint64_t parent, uint16_t m) {                                       /* :170-172 */
  if (o->overflow) return;
  o->buf[parent + 5] = (uint8_t)(m >> 8); o->buf[parent + 6] = (uint8_t)m;
}

/* :620-649 */
static int gen_surface_normal(const oct_t *o, const int c[3], int16_t *packed) {
  int exposed = 0, nx = 0, ny = 0, nz = 0;
  for (int i = c[0] - 1; i <= c[0] + 1; i++) {
    if (i < 0 || i >= o->chunk) continue;
    for (int j = c[1] - 1; j <= c[1] + 1; j++) {
      if (j < 0 || j >= o->chunk) continue;
      for (int k = c[2] - 1; k <= c[2] + 1; k++) {
        if (k < 0 || k >= o->chunk) continue;
        if (get_voxel(o, i, j, k) == 0) { exposed = 1; nx += i - c[0]; ny += j - c[1]; nz += k - c[2]; }
      }
    }
  }
  nx = nx / 2 + 5; ny = ny / 2 + 5; nz = nz / 2 + 5; /* Java int division truncates toward zero, as C */
  *packed = (int16_t)(nx + ny * 10 + nz * 100);
  return exposed;
}

/* :651-670 -- probes only {c-1, c+s, c+s+1}^3 */
static int check_big_node_exposed(const oct_t *o, const int c[3], int s) {
  int exposed = 0;
  for (int i = c[2] - 1; i <= c[2] + s + 1; i++) {
    if (i < 0 || i >= o->chunk || (i >= c[2] && i <= c[2] + s - 1)) continue;
    for (int j = c[1] - 1; j <= c[1] + s + 1; j++) {
      if (j < 0 || j >= o->chunk || (j >= c[1] && j <= c[1] + s - 1)) continue;
      for (int k = c[0] - 1; k <= c[0] + s + 1; k++) {
        if (k < 0 || k >= o->chunk || (k >= c[0] && k <= c[0] + s - 1)) continue;
        if (get_voxel(o, k, j, i) == 0) exposed = 1;
      }
    }
  }
  return exposed;
}

enum { T_INTERIOR = 0, T_SURFACE = 1, T_SUBDIV = 2, T_NONSURF = 3 };

/* :511-608 */
static void construct_inner(oct_t *o, int size, int curLOD, int maxLOD, const int pPos[3], uint64_t parent) {
  int cSize = size / 2;
  if (cSize == 0 || curLOD == maxLOD) return;
  uint64_t children[8];
  int types[8];
  int cPos[8][3];
  for (int n = 0; n < 8; n++)
    for (int a = 0; a < 3; a++) cPos[n][a] = pPos[a] + child_offsets[n][a] * cSize;
  uint16_t leafMask = 0;
  for (int n = 0; n < 8; n++) {
    uint8_t first = get_voxel(o, cPos[n][0], cPos[n][1], cPos[n][2]);
    uint8_t value = first;
    int leaf = 1, type = T_INTERIOR;
    if (curLOD + 1 != maxLOD) {
      for (int i = cPos[n][2]; i < cPos[n][2] + cSize && leaf; i++)
        for (int j = cPos[n][1]; j < cPos[n][1] + cSize && leaf; j++)
          for (int k = cPos[n][0]; k < cPos[n][0] + cSize; k++) {
            uint8_t sample = get_voxel(o, k, j, i);
            if (sample != 0) value = sample;
            if (sample != first) {
              if (first == 0) first = sample;
              value = first;
              leaf = 0;
              break;
            }
          }
    }
    if (leaf && value != 0) {
      if (cSize == 1) {
        int16_t normal;
        if (gen_surface_normal(o, cPos[n], &normal)) { children[n] = create_surface(o, value, normal); type = T_SURFACE; }
        else { children[n] = create_non_surface(o, value); type = T_NONSURF; }
      } else {
        if (check_big_node_exposed(o, cPos[n], cSize)) { leaf = 0; children[n] = create_interior(o, value); type = T_INTERIOR; }
        else { children[n] = create_subdividable(o, value); type = T_SUBDIV; }
      }
    } else if (leaf) {
      if (cSize == 1) { children[n] = create_non_surface(o, value); type = T_NONSURF; }
      else { children[n] = create_subdividable(o, value); type = T_SUBDIV; }
    } else {
      children[n] = create_interior(o, value);
      type = T_INTERIOR;
    }
    types[n] = type;
    leafMask |= (uint16_t)(type << (n << 1)); /* :589-599 */
  }
  if (o->overflow) return;
  set_child_pointer(o, parent, children[0]);
  set_leaf_mask(o, parent, leafMask);
  for (int n = 0; n < 8; n++)
    if (o->buf[children[n]] != 0 && types[n] == T_INTERIOR)
      construct_inner(o, cSize, curLOD + 1, maxLOD, cPos[n], children[n]);
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

uint64_t svo_oracle_build_dense(const uint8_t *voxels, int n, uint8_t *out, uint64_t cap, uint64_t counts[4]) {
  oct_t o;
  memset(&o, 0, sizeof o);
  o.buf = out; o.cap = cap; o.vox = voxels; o.lg = ilog2(n); o.chunk = n;
  int zero[3] = {0, 0, 0};
  create_interior(&o, 1); /* createDummyHead, OctreeThread.java:21 */
  construct_inner(&o, n, 0, ilog2(n), zero, 0);
  if (counts) memcpy(counts, o.counts, sizeof o.counts);
  return o.overflow ? 0 : o.off;
}

typedef struct { int origin[3]; uint64_t pointer; } chunk_t;

/* :481-502 */
static void fill_empty_children(oct_t *o, uint64_t parent, int levels, const int pPos[3], int chunk,
                                chunk_t *chunks, int *nchunks) {
  if (levels == 0) {
    chunk_t *c = &chunks[(*nchunks)++];
    memcpy(c->origin, pPos, sizeof c->origin);
    c->pointer = parent;
    return;
  }
  int cSize = chunk << (levels - 1);
  uint64_t children[8];
  int cPos[8][3];
  for (int n = 0; n < 8; n++)
    for (int a = 0; a < 3; a++) cPos[n][a] = pPos[a] + child_offsets[n][a] * cSize;
  for (int i = 0; i < 8; i++) children[i] = create_interior(o, 1);
  for (int i = 0; i < 8; i++) fill_empty_children(o, children[i], levels - 1, cPos[i], chunk, chunks, nchunks);
  set_child_pointer(o, parent, children[0]);
}

uint64_t svo_oracle_build_terrain(const uint16_t *height, const uint8_t *mat, int n, int chunk,
                                  uint8_t *out, uint64_t cap, uint64_t counts[4]) {
  if (chunk > n) chunk = n;
  int levels = ilog2(n / chunk);
  int lg = ilog2(chunk);
  oct_t o;
  memset(&o, 0, sizeof o);
  o.buf = out; o.cap = cap; o.lg = lg; o.chunk = chunk;
  memset(out, 0, cap); /* BufferUtils.createByteBuffer zero-fills (Octree.java:65) */

  int nchunks_max = 1;
  for (int l = 0; l < levels; l++) nchunks_max *= 8;
  chunk_t *chunks = (chunk_t *)calloc((size_t)nchunks_max, sizeof *chunks);
  int nchunks = 0;
  int rootPos[3] = {0, 0, 0};
  create_interior(&o, 1);                                              /* :234 */
  fill_empty_children(&o, 0, levels, rootPos, chunk, chunks, &nchunks); /* :244 */

  size_t cvox = (size_t)chunk * chunk * chunk;
  uint8_t *vox = (uint8_t *)malloc(cvox);
  uint64_t subcap = cap;
  uint8_t *sub = (uint8_t *)malloc(subcap);
  for (int ci = 0; ci < nchunks && !o.overflow; ci++) {
    const chunk_t *c = &chunks[ci];
    /* chunkgen-heightmap.comp:13-31 */
    for (int z = 0; z < chunk; z++)
      for (int x = 0; x < chunk; x++) {
        size_t hm = (size_t)(c->origin[2] + z) * (size_t)n + (size_t)(c->origin[0] + x);
        int hs = (int)(((uint32_t)height[hm] * (uint32_t)(n / 4)) >> 16);
        uint8_t ms = mat[hm];
        for (int y = 0; y < chunk; y++) {
          int posY = y + c->origin[1];
          uint8_t v = 0;
          if (posY <= hs) v = (hs - posY <= 4) ? ms : 1;
          vox[(size_t)x | ((size_t)y << lg) | ((size_t)z << (2 * lg))] = v;
        }
      }
    /* :292-343: 8 OctreeThreads, then splice */
    int cSize = chunk / 2;
    uint64_t children[8];
    uint64_t sub_len[8];
    uint16_t sub_mask[8];
    uint8_t *sub_bufs[8];
    for (int i = 0; i < 8; i++) {
      oct_t t;
      memset(&t, 0, sizeof t);
      t.buf = sub; t.cap = subcap; t.vox = vox; t.lg = lg; t.chunk = chunk;
      int cp[3] = {child_offsets[i][0] * cSize, child_offsets[i][1] * cSize, child_offsets[i][2] * cSize};
      create_interior(&t, 1);                                 /* createDummyHead */
      construct_inner(&t, cSize, 0, ilog2(cSize), cp, 0);     /* OctreeThread.run: (512, 0, 9) upstream */
      if (t.overflow) { o.overflow = 1; break; }
      sub_len[i] = t.off;
      sub_mask[i] = (uint16_t)((sub[5] << 8) | sub[6]);
      sub_bufs[i] = (uint8_t *)malloc(t.off);
      memcpy(sub_bufs[i], sub, t.off);
      for (int k = 0; k < 4; k++) o.counts[k] += t.counts[k];
    }
    if (o.overflow) break;
    for (int i = 0; i < 8; i++) children[i] = create_interior(&o, 1);  /* :317-319 */
    set_child_pointer(&o, c->pointer, children[0]);                    /* :320 */
    for (int i = 0; i < 8; i++) {
      if (o.off + sub_len[i] > o.cap) { o.overflow = 1; }
      if (!o.overflow) {
        set_child_pointer(&o, children[i], o.off);                     /* :330 */
        set_leaf_mask(&o, children[i], sub_mask[i]);                   /* :331 */
        memcpy(o.buf + o.off, sub_bufs[i] + 7, sub_len[i] - 7);        /* :334-335 */
        o.off += sub_len[i];                                           /* :336 (advances 7 past the copy) */
      }
      free(sub_bufs[i]);
    }
  }
  free(sub);
  free(vox);
  free(chunks);
  if (counts) memcpy(counts, o.counts, sizeof o.counts);
  return o.overflow ? 0 : o.off;
}
