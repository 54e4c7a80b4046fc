// svo_transcode.h -- node-stream -> descriptor transcode done at upload time.
//
// Input: the engine's node buffer exactly as Octree.java writes it
// (src/engine/Octree.java:68-95,119-176): interior = [value u8][childPtr i32
// big-endian, relative to the node's own offset][leafMask u16 big-endian, one
// 2-bit type code per child]; surface leaf = [value][normal u16 little-endian];
// non-surface leaf = [value]; subdividable leaf = [value][6 x 0].
//
// Output (see svo_trace.cuh): one uint2 descriptor per interior node that the
// traversal can descend into, in breadth-first order, plus the reference byte
// offset of each node's child block (for hit ids and hit-record decoding).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>  // uint2

namespace svo {

// Axis-aligned box of octree cells in integer cell units of depth 24 (cube edge = 2^24), empty if lo > hi.
struct CellBox {
  uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  uint32_t hi[3] = {0u, 0u, 0u};
  bool empty() const { return lo[0] > hi[0]; }
  void add(const CellBox &o) {
    for (int a = 0; a < 3; a++) {
      if (o.lo[a] < lo[a]) lo[a] = o.lo[a];
      if (o.hi[a] > hi[a]) hi[a] = o.hi[a];
    }
  }
};

struct Transcoded {
  std::vector<uint2> desc;
  std::vector<uint32_t> refbase;
  std::vector<uint2> meta;            // per node: .x = byte offset of its own record, .y = index of its parent (root: 0xFFFFFFFF);
                                      // what the incremental transcode (svo_upload_range) needs to find the nodes an edit touches
  std::vector<uint32_t> level_start;  // desc index where each BFS level begins
  // Where a cast can end in a hit: `leaf_box` bounds every record with value != 0 that the traversal treats
  // as a leaf (child.cp == 0, svotrace.comp:311), `depth_box[d]` bounds every record with value != 0 at tree
  // depth d (a cast stops there when d == maxDepth, svotrace.comp:300).  Rays outside the union of the boxes
  // that apply to a frame cannot hit anything.
  CellBox leaf_box;
  CellBox depth_box[24];
};

// Content box of a frame in cube coordinates [1,2], padded by 2^-9 (see Trav::setup): leaves at any depth, plus
// every non-empty record at the depths where the frame's casts stop (maxDepth, and coneDepth for the cone-traced
// bounces); depths < 0 = all depths.  An empty box comes out as lo > hi.
inline void content_box(const CellBox &leaf_box, const CellBox *depth_box, int maxDepth, int coneDepth, float lo[3], float hi[3]) {
  CellBox b = leaf_box;
  if (maxDepth >= 0 && maxDepth < 24 && coneDepth >= 0 && coneDepth < 24) {
    b.add(depth_box[maxDepth]);
    b.add(depth_box[coneDepth]);
  } else {
    for (int d = 0; d < 24; d++) b.add(depth_box[d]);
  }
  for (int a = 0; a < 3; a++) {
    if (b.empty()) { lo[a] = 4.0f; hi[a] = -4.0f; continue; }
    lo[a] = 1.0f + (float)b.lo[a] * (1.0f / 16777216.0f) - 0.001953125f;
    hi[a] = 1.0f + (float)b.hi[a] * (1.0f / 16777216.0f) + 0.001953125f;
  }
}

// Returns false (with `err` set) if the stream cannot be a tree (more
// descriptors than bytes: cyclic or heavily aliased child pointers).
bool transcode_stream(const uint8_t *raw, uint64_t nbytes, Transcoded &out, std::string &err, int nthreads = 0);

}  // namespace svo
