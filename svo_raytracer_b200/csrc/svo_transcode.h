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

struct Transcoded {
  std::vector<uint2> desc;
  std::vector<uint32_t> refbase;
  std::vector<uint32_t> level_start;  // desc index where each BFS level begins
};

// Returns false (with `err` set) if the stream cannot be a tree (more
// descriptors than bytes: cyclic or heavily aliased child pointers).
bool transcode_stream(const uint8_t *raw, uint64_t nbytes, Transcoded &out, std::string &err);

}  // namespace svo
