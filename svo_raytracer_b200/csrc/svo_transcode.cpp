// svo_transcode.cpp -- upload-time transcode (see svo_transcode.h): level-synchronous breadth-first walk of the
// reference node stream, each level processed by all host threads in two passes (count the descriptor-bearing
// children of every node; exclusive scan; emit descriptors and the next level's work list at their final slots),
// so the output is identical for any thread count.
#include "svo_transcode.h"

#include <algorithm>
#include <thread>

namespace svo {
namespace {

struct Pending {
  uint32_t off;      // byte offset of the node record (Node.descriptor, svotrace.comp:85)
  uint32_t cp;       // its childPtr (relative)
  uint32_t codes;    // its leafMask
  uint32_t x, y, z;  // cell coordinates at the node's own depth
};

inline uint32_t rd(const uint8_t *raw, uint64_t n, uint32_t p) { return ((uint64_t)p < n) ? raw[p] : 0u; }  // getByte, out of range = 0
inline uint32_t rd_be32(const uint8_t *raw, uint64_t n, uint32_t p) {
  return (rd(raw, n, p) << 24) | (rd(raw, n, p + 1u) << 16) | (rd(raw, n, p + 2u) << 8) | rd(raw, n, p + 3u);
}
inline uint32_t rd_be16(const uint8_t *raw, uint64_t n, uint32_t p) { return (rd(raw, n, p) << 8) | rd(raw, n, p + 1u); }

// One node of the current level.  `emit` == nullptr: only count its descriptor-bearing children.
struct NodeOut {
  uint32_t ndesc_children;
};

inline uint32_t visit(const uint8_t *raw, uint64_t nbytes, const Pending &nd, int depth, uint32_t child_base, Pending *next_out,
                      uint2 *desc_out, uint32_t *refbase_out, CellBox *leaf_box, CellBox *depth_box, uint2 *meta_out = nullptr,
                      uint32_t my_index = 0) {
  const uint32_t ref_base = nd.off + nd.cp;  // uint wrap-around as in extractChild (:134)
  uint32_t p = ref_base, nonzero = 0, has_desc = 0, n_next = 0;
  for (uint32_t c = 0; c < 8; c++) {
    const uint32_t code = (nd.codes >> (2u * c)) & 3u;
    const uint32_t size = code == 1u ? 3u : (code == 3u ? 1u : 7u);
    const uint32_t value = rd(raw, nbytes, p);
    if (value != 0u) {
      nonzero |= 1u << c;
      const uint32_t ccp = code == 0u ? rd_be32(raw, nbytes, p + 1u) : 0u;
      const uint32_t cx = 2u * nd.x + (c & 1u), cy = 2u * nd.y + ((c >> 1) & 1u), cz = 2u * nd.z + ((c >> 2) & 1u);
      if (desc_out) {  // emit pass: bounds of where casts can end
        const int sh = 24 - (depth + 1);  // child cells are at tree depth depth+1 <= 23
        CellBox cell;
        cell.lo[0] = cx << sh; cell.lo[1] = cy << sh; cell.lo[2] = cz << sh;
        cell.hi[0] = (cx + 1u) << sh; cell.hi[1] = (cy + 1u) << sh; cell.hi[2] = (cz + 1u) << sh;
        depth_box->add(cell);
        if (ccp == 0u) leaf_box->add(cell);  // child.cp == 0: a hit wherever the traversal meets it (:311)
      }
      if (ccp != 0u && depth < 22) {  // child.cp != 0: the traversal may PUSH into it
        has_desc |= 1u << c;
        if (next_out) next_out[n_next] = {p, ccp, rd_be16(raw, nbytes, p + 5u), cx, cy, cz};
        if (meta_out) meta_out[n_next] = make_uint2(p, my_index);
        n_next++;
      }
    }
    p += size;
  }
  if (desc_out) {
    desc_out->x = child_base;
    desc_out->y = (nd.codes & 0xFFFFu) | (nonzero << 16) | (has_desc << 24);
    *refbase_out = ref_base;
  }
  return n_next;
}

template <class F>
void parallel_chunks(size_t n, int nthreads, F f) {  // f(chunk_index, begin, end)
  const size_t nchunks = (size_t)nthreads;
  if (nthreads <= 1 || n < 4096) {
    f(0, 0, n);
    return;
  }
  std::vector<std::thread> pool;
  for (size_t t = 0; t < nchunks; t++) {
    const size_t b = n * t / nchunks, e = n * (t + 1) / nchunks;
    pool.emplace_back(f, t, b, e);
  }
  for (auto &th : pool) th.join();
}

}  // namespace

bool transcode_stream(const uint8_t *raw, uint64_t nbytes, Transcoded &out, std::string &err, int nthreads) {
  out.desc.clear();
  out.refbase.clear();
  out.meta.clear();
  out.meta.push_back(make_uint2(0u, 0xFFFFFFFFu));
  out.level_start.clear();
  out.leaf_box = CellBox();
  for (CellBox &b : out.depth_box) b = CellBox();
  if (nbytes >= (1ull << 32)) {
    err = "node stream must be < 4 GiB (the engine addresses it with int32 byte offsets)";
    return false;
  }
  if (nthreads < 1) nthreads = (int)std::thread::hardware_concurrency();
  nthreads = std::max(1, std::min(nthreads, 64));
  const uint64_t limit = (nbytes > 4096 ? nbytes : 4096);
  std::vector<Pending> cur, next;
  cur.push_back({0u, rd_be32(raw, nbytes, 1u), rd_be16(raw, nbytes, 5u), 0u, 0u, 0u});  // extractNode(0), svotrace.comp:222
  std::vector<uint32_t> counts;
  // Tree depth D nodes have children at scale 22-D; scale 0 children (D = 22)
  // can only be hit, never entered (maxDepth <= 23), so 23 levels suffice.
  for (int depth = 0; depth <= 22 && !cur.empty(); depth++) {
    const size_t n = cur.size();
    const uint64_t level_base = out.desc.size();
    out.level_start.push_back((uint32_t)level_base);
    // pass 1: descriptor-bearing children per node
    counts.resize(n + 1);
    parallel_chunks(n, nthreads, [&](size_t, size_t b, size_t e) {
      for (size_t i = b; i < e; i++) counts[i] = visit(raw, nbytes, cur[i], depth, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    });
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) { const uint32_t c = counts[i]; counts[i] = (uint32_t)total; total += c; }
    if (level_base + n + total > limit) {
      err = "node stream is not a tree (child pointers alias or cycle)";
      return false;
    }
    // pass 2: emit
    out.desc.resize(level_base + n);
    out.refbase.resize(level_base + n);
    next.resize(total);
    const uint64_t next_base = level_base + n;
    out.meta.resize(next_base + total);
    std::vector<CellBox> leaf_boxes((size_t)nthreads), depth_boxes((size_t)nthreads);
    parallel_chunks(n, nthreads, [&](size_t t, size_t b, size_t e) {
      for (size_t i = b; i < e; i++)
        visit(raw, nbytes, cur[i], depth, (uint32_t)(next_base + counts[i]), next.data() + counts[i], &out.desc[level_base + i],
              &out.refbase[level_base + i], &leaf_boxes[t], &depth_boxes[t], out.meta.data() + next_base + counts[i],
              (uint32_t)(level_base + i));
    });
    for (int t = 0; t < nthreads; t++) {
      out.leaf_box.add(leaf_boxes[(size_t)t]);
      out.depth_box[depth + 1].add(depth_boxes[(size_t)t]);
    }
    cur.swap(next);
  }
  return true;
}

}  // namespace svo
