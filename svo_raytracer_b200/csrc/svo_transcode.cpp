// svo_transcode.cpp -- host BFS transcode (see svo_transcode.h).
#include "svo_transcode.h"

namespace svo {
namespace {

struct Pending {
  uint32_t off;    // byte offset of the node record (Node.descriptor, svotrace.comp:85)
  uint32_t cp;     // its childPtr (relative)
  uint32_t codes;  // its leafMask
  uint32_t x, y, z;  // cell coordinates at the node's own depth
};

inline uint32_t rd(const uint8_t *raw, uint64_t n, uint32_t p) { return ((uint64_t)p < n) ? raw[p] : 0u; }  // getByte, out of range = 0
inline uint32_t rd_be32(const uint8_t *raw, uint64_t n, uint32_t p) {
  return (rd(raw, n, p) << 24) | (rd(raw, n, p + 1u) << 16) | (rd(raw, n, p + 2u) << 8) | rd(raw, n, p + 3u);
}
inline uint32_t rd_be16(const uint8_t *raw, uint64_t n, uint32_t p) { return (rd(raw, n, p) << 8) | rd(raw, n, p + 1u); }

}  // namespace

bool transcode_stream(const uint8_t *raw, uint64_t nbytes, Transcoded &out, std::string &err) {
  out.desc.clear();
  out.refbase.clear();
  out.level_start.clear();
  out.leaf_box = CellBox();
  for (CellBox &b : out.depth_box) b = CellBox();
  if (nbytes >= (1ull << 32)) {
    err = "node stream must be < 4 GiB (the engine addresses it with int32 byte offsets)";
    return false;
  }
  const uint64_t limit = (nbytes > 4096 ? nbytes : 4096);
  std::vector<Pending> cur, next;
  cur.push_back({0u, rd_be32(raw, nbytes, 1u), rd_be16(raw, nbytes, 5u), 0u, 0u, 0u});  // extractNode(0), svotrace.comp:222
  // Tree depth D nodes have children at scale 22-D; scale 0 children (D = 22)
  // can only be hit, never entered (maxDepth <= 23), so 23 levels suffice.
  for (int depth = 0; depth <= 22 && !cur.empty(); depth++) {
    out.level_start.push_back((uint32_t)out.desc.size());
    const uint64_t next_base = out.desc.size() + cur.size();
    next.clear();
    for (const Pending &nd : cur) {
      const uint32_t ref_base = nd.off + nd.cp;  // uint wrap-around as in extractChild (:134)
      uint32_t p = ref_base, nonzero = 0, has_desc = 0;
      const uint32_t child_base = (uint32_t)(next_base + next.size());
      for (uint32_t c = 0; c < 8; c++) {
        const uint32_t code = (nd.codes >> (2u * c)) & 3u;
        const uint32_t size = code == 1u ? 3u : (code == 3u ? 1u : 7u);
        const uint32_t value = rd(raw, nbytes, p);
        if (value != 0u) {
          nonzero |= 1u << c;
          const uint32_t cx = 2u * nd.x + (c & 1u), cy = 2u * nd.y + ((c >> 1) & 1u), cz = 2u * nd.z + ((c >> 2) & 1u);
          const int cdepth = depth + 1, sh = 24 - cdepth;  // child cells are at tree depth depth+1 <= 23
          CellBox cell;
          cell.lo[0] = cx << sh; cell.lo[1] = cy << sh; cell.lo[2] = cz << sh;
          cell.hi[0] = (cx + 1u) << sh; cell.hi[1] = (cy + 1u) << sh; cell.hi[2] = (cz + 1u) << sh;
          out.depth_box[cdepth].add(cell);
          const uint32_t ccp = code == 0u ? rd_be32(raw, nbytes, p + 1u) : 0u;
          if (ccp == 0u) out.leaf_box.add(cell);  // child.cp == 0: a hit wherever the traversal meets it (:311)
          else if (depth < 22) {                  // child.cp != 0: the traversal may PUSH into it
            has_desc |= 1u << c;
            next.push_back({p, ccp, rd_be16(raw, nbytes, p + 5u), cx, cy, cz});
          }
        }
        p += size;
      }
      uint2 d;
      d.x = child_base;
      d.y = (nd.codes & 0xFFFFu) | (nonzero << 16) | (has_desc << 24);
      out.desc.push_back(d);
      out.refbase.push_back(ref_base);
    }
    if (out.desc.size() + next.size() > limit) {
      err = "node stream is not a tree (child pointers alias or cycle)";
      return false;
    }
    cur.swap(next);
  }
  return true;
}

}  // namespace svo
