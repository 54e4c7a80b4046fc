// ref_java_harness.cpp -- TEST INFRASTRUCTURE.  extern "C" entry points around the reference's octree builder and SDF
// brush as compiled from their own Java text (oracle/build_ref_java.py -> oracle/_ref/ref_java_gen.inc).  What this file
// restates is only what surrounds those methods in the engine:
//   * Octree.constructCompleteOctree's frame (Octree.java:244-262): root = createInteriorNode(1), fillEmptyChildren(0,
//     chunkLevel, rootPos, chunks), then per chunk the voxel volume and the (extracted) thread-and-splice block;
//   * the dispatch of the voxeliser per chunk (Octree.java:270-282): the shader itself, chunkgen-heightmap.comp, is compiled
//     from its text too (oracle/glsl_voxel_shim.h) and run per voxel; its image is read back by the engine into a ByteBuffer
//     indexed x | y << 10 | z << 20 (Octree.java:110-112), which is where the invocations store here;
//   * Main.placeSDF (Main.java:338-353): `new Sphere(pos, r)`, `octree.useSDFBrush(sdf, value)`, the two ranges.
// Settable constants (defaults = shipped): Octree::CHUNK_SIZE, Constants::SUB_OCTREE_SIZE, Constants::WORLD_SIZE,
// Constants::SDF_MAX_LOD -- see rule J9 of the generator.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "glsl_voxel_shim.h"
#include "java_shim.h"

namespace javaref {
#include "ref_java_gen.inc"
}  // namespace javaref
#include "ref_chunkgen_gen.inc"  // namespace glslv::ref_chunkgen: chunkgen-heightmap.comp

using namespace javaref;

namespace {
int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return l;
}
struct Saved {
  int chunk = Octree::CHUNK_SIZE, sub = Constants::SUB_OCTREE_SIZE, world = Constants::WORLD_SIZE, lod = Constants::SDF_MAX_LOD,
      subkb = Constants::SUB_OCTREE_MEMORY_SIZE_KB;
  size_t mark = BufferUtils::mark();
  ~Saved() {
    Octree::CHUNK_SIZE = chunk;
    Constants::SUB_OCTREE_SIZE = sub;
    Constants::WORLD_SIZE = world;
    Constants::SDF_MAX_LOD = lod;
    Constants::SUB_OCTREE_MEMORY_SIZE_KB = subkb;
    BufferUtils::release_to(mark);
  }
};
int kb_for(uint64_t bytes) { return (int)((bytes + 1023) / 1024) + 1; }
}  // namespace

extern "C" {

const char *svo_refj_about(void) {
  return "reference octree builder + SDF brush: Octree.java / OctreeThread.java / Util.java / sdf/*.java rewritten mechanically "
         "(oracle/build_ref_java.py) and compiled by g++";
}

// One OctreeThread (OctreeThread.java:19-23) over a dense n^3 volume (voxels[x | y << lg | z << 2 lg], lg = log2 n): dummy head +
// constructInnerOctree(n, 0, log2 n, {0,0,0}, 0, voxelBuffer), with CHUNK_SIZE = n for the neighbourhood tests.  n <= 1024.
uint64_t svo_refj_build_dense(const uint8_t *voxels, int n, uint8_t *out, uint64_t cap, uint64_t counts[4]) try {
  if (n < 1 || n > 1024 || (n & (n - 1))) return 0;
  Saved saved;
  const int lg = ilog2(n);
  Octree::CHUNK_SIZE = n;
  ByteBuffer *vox = BufferUtils::createByteBuffer(1 << 30);
  for (int z = 0; z < n; z++)
    for (int y = 0; y < n; y++)
      memcpy(vox->data + ((size_t)y << 10 | (size_t)z << 20), voxels + ((size_t)y << lg | (size_t)z << (2 * lg)), (size_t)n);
  Octree *o = new Octree(kb_for(cap));
  o->createDummyHead();
  jarray<int> zero = jarray<int>{0, 0, 0};
  o->constructInnerOctree(n, 0, lg, zero, 0, vox);
  const uint64_t len = (uint64_t)o->memOffset;
  if (counts) {
    counts[0] = (uint64_t)o->surfaceLeafNodes;
    counts[1] = (uint64_t)o->nonSurfaceLeafNodes;
    counts[2] = (uint64_t)o->subdividableLeafNodes;
    counts[3] = (uint64_t)o->interiorNodes;
  }
  if (len <= cap) memcpy(out, o->buffer->data, len);
  delete o;
  return len <= cap ? len : 0;
} catch (const std::exception &) {
  return 0;  // the buffer was too small (Java: IndexOutOfBoundsException)
}

// constructCompleteOctree for a heightmap world of edge n = chunk << levels (upstream: chunk 1024, levels 3, n = 8192).
// height: n*n u16, mat: n*n u8, both [z*n + x]; column height in voxels = (height * (n / 4)) >> 16.
uint64_t svo_refj_build_terrain(const uint16_t *height, const uint8_t *mat, int n, int chunk, uint8_t *out, uint64_t cap, uint64_t counts[4]) try {
  if (chunk > n) chunk = n;
  if (chunk < 2 || chunk > 1024 || (chunk & (chunk - 1)) || (n & (n - 1))) return 0;
  Saved saved;
  Octree::CHUNK_SIZE = chunk;
  Constants::SUB_OCTREE_SIZE = chunk / 2;
  Constants::SUB_OCTREE_MEMORY_SIZE_KB = kb_for(cap);
  const int levels = ilog2(n / chunk);
  Octree *o = new Octree(kb_for(cap));
  jarray<int> rootPos = jarray<int>{0, 0, 0};
  o->createInteriorNode((jbyte)1);                                   // Octree.java:247
  ArrayList<Octree::Chunk *> *chunks = new ArrayList<Octree::Chunk *>();
  o->fillEmptyChildren(0, levels, rootPos, chunks);                  // :257
  ByteBuffer *vox = BufferUtils::createByteBuffer(1 << 30);          // :296 (CHUNK_SIZE^3 upstream: the 1024 pitch is getVoxel's)
  {
    namespace cg = glslv::ref_chunkgen;
    cg::ref_height_scale = n / 4;                                    // the shader's literal 2048 = 8192 / 4
    cg::heightImage.texels = height; cg::heightImage.width = cg::heightImage.height = n;          // image unit 4 (Octree.java:214)
    cg::matImage.texels = (const int8_t *)mat; cg::matImage.width = cg::matImage.height = n;      // image unit 5 (:226)
    cg::voxelImage.texels = (int8_t *)vox->data;                                                  // image unit 3
    cg::voxelImage.width = cg::voxelImage.height = cg::voxelImage.depth = chunk;
    cg::voxelImage.row_pitch = 1u << 10; cg::voxelImage.slice_pitch = 1u << 20;
  }
  for (Octree::Chunk *c : *chunks) {
    const size_t mark = BufferUtils::mark();
    // Octree.java:270-282: offsets = chunk origin, glDispatchCompute over the chunk, glGetTexImage into the voxel buffer.
    // The dispatch runs the reference's voxeliser itself (chunkgen-heightmap.comp compiled from its text), one invocation
    // per voxel, storing into the 1024-pitch layout getVoxel reads.
    namespace cg = glslv::ref_chunkgen;
    cg::offsetX = c->origin[0];
    cg::offsetY = c->origin[1];
    cg::offsetZ = c->origin[2];
    cg::Invocation inv;
    for (int z = 0; z < chunk; z++)
      for (int y = 0; y < chunk; y++)
        for (int x = 0; x < chunk; x++) {
          inv.gl_GlobalInvocationID.set((uint32_t)x, (uint32_t)y, (uint32_t)z);
          inv.main();
        }
    o->buildChunk(c, vox, ilog2(chunk / 2));                          // :285-338 (maxLOD = 9 upstream)
    BufferUtils::release_to(mark);                                    // the eight sub-octree buffers
  }
  const uint64_t len = (uint64_t)o->memOffset;
  if (counts) {
    counts[0] = (uint64_t)o->surfaceLeafNodes;
    counts[1] = (uint64_t)o->nonSurfaceLeafNodes;
    counts[2] = (uint64_t)o->subdividableLeafNodes;
    counts[3] = (uint64_t)o->interiorNodes;
  }
  if (len <= cap) memcpy(out, o->buffer->data, len);
  delete chunks;
  delete o;
  return len <= cap ? len : 0;
} catch (const std::exception &) {
  return 0;
}

// Main.placeSDF on a stream of *nbytes bytes held in a buffer of cap bytes: kind 0 = Sphere(origin, p[0]) (the engine's brush),
// 1 = Box(origin, p[0], p[1], p[2]); value 0 subtracts.  The stream is edited in place; bounds = {start0, end0, start1, end1}
// (Octree.ChangeBounds), *nbytes = the new memOffset.  world_size / max_lod: Constants.WORLD_SIZE (8196 upstream -- sic) and the
// hard-coded 13 of Octree.java:706.
int svo_refj_sdf_brush(uint8_t *stream, uint64_t *nbytes, uint64_t cap, int world_size, int max_lod, int kind, const int origin[3],
                       const int p[3], int value, int64_t bounds[4]) try {
  if (!stream || !nbytes || *nbytes > cap || cap > 0x7fff0000ull) return 1;
  Saved saved;
  Constants::WORLD_SIZE = world_size;
  Constants::SDF_MAX_LOD = max_lod;
  Octree *o = new Octree(kb_for(cap));
  memcpy(o->buffer->data, stream, *nbytes);
  o->memOffset = (int)*nbytes;
  jarray<int> org = jarray<int>{origin[0], origin[1], origin[2]};
  SignedDistanceField *sdf = kind == 0 ? (SignedDistanceField *)new Sphere(org, p[0]) : (SignedDistanceField *)new Box(org, p[0], p[1], p[2]);
  Octree::ChangeBounds *cb = o->useSDFBrush(sdf, (jbyte)value);      // Main.java:345
  bounds[0] = cb->start0;
  bounds[1] = cb->end0;
  bounds[2] = cb->start1;
  bounds[3] = cb->end1;
  int rc = 0;
  if ((uint64_t)o->memOffset > cap) rc = 2;
  else {
    memcpy(stream, o->buffer->data, (size_t)o->memOffset);
    *nbytes = (uint64_t)o->memOffset;
  }
  delete cb;
  delete o;  // (the brush object is left to the process, like everything the Java text makes with `new`)
  return rc;
} catch (const std::exception &) {
  return 3;
}

}  // extern "C"
