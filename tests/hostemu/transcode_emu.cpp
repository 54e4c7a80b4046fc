// transcode_emu.cpp -- TEST INFRASTRUCTURE: svo_gpu_transcode.cu (whole and incremental upload-time transcode) compiled by
// g++ and run on the coroutine SIMT emulator, against the host transcode (svo_transcode.cpp).
#include "cuda_host_shim.h"
#include "simt_emu.h"

#define SVO_HOST_EMU 1
#include "../../svo_raytracer_b200/csrc/svo_gpu_transcode.cu"

#include <string>
#include <utility>
#include <vector>

using namespace svo;

namespace {
// layout-independent fingerprint (the same walk as svo_scene_canonical in svo_capi.cu)
bool canonical(const std::vector<uint2> &desc, const std::vector<uint32_t> &ref, uint64_t nd, uint64_t out[3]) {
  uint64_t h = 1469598103934665603ull, reachable = 0, max_depth = 0;
  auto mix = [&](uint32_t v) { for (int i = 0; i < 4; i++) { h ^= (v >> (8 * i)) & 0xFFu; h *= 1099511628211ull; } };
  std::vector<std::pair<uint32_t, uint32_t>> stack;
  if (nd) stack.push_back({0u, 0u});
  while (!stack.empty()) {
    const uint32_t i = stack.back().first, d = stack.back().second;
    stack.pop_back();
    if (i >= nd || reachable > nd) return false;
    reachable++;
    if (d > max_depth) max_depth = d;
    mix(desc[i].y); mix(ref[i]); mix(d);
    const uint32_t has = desc[i].y >> 24;
    uint32_t k = (uint32_t)__builtin_popcount(has);
    for (int c = 7; c >= 0; c--)
      if ((has >> c) & 1u) stack.push_back({desc[i].x + --k, d + 1u});
  }
  out[0] = reachable; out[1] = h; out[2] = max_depth;
  return true;
}
bool box_contains(const CellBox &outer, const CellBox &inner) {
  if (inner.empty()) return true;
  for (int a = 0; a < 3; a++)
    if (outer.lo[a] > inner.lo[a] || outer.hi[a] < inner.hi[a]) return false;
  return true;
}
}  // namespace

extern "C" {

// whole transcode on the emulator against the host transcode: 0 = identical arrays (desc, refbase, meta) and boxes
int emu_gpu_transcode_check(const uint8_t *raw, uint64_t nbytes, int nthreads) {
  Transcoded t;
  std::string err;
  if (!transcode_stream(raw, nbytes, t, err, 2)) return 1;
  const uint64_t cap = nbytes / 7 + 4096;
  std::vector<uint2> desc(cap), meta(cap);
  std::vector<uint32_t> ref(cap);
  uint64_t nd = 0;
  uint32_t nlevels = 0;
  CellBox lb, db[24];
  bool overflow = false;
  simt::g_os_threads = nthreads;
  if (gpu_transcode(raw, nbytes, desc.data(), ref.data(), meta.data(), cap, &nd, &nlevels, &lb, db, &overflow, nullptr) != cudaSuccess) return 2;
  if (overflow) return 3;
  if (nd != t.desc.size() || nlevels != t.level_start.size()) return 4;
  for (uint64_t i = 0; i < nd; i++) {
    if (desc[i].x != t.desc[i].x || desc[i].y != t.desc[i].y || ref[i] != t.refbase[i]) return 5;
    if (meta[i].x != t.meta[i].x || meta[i].y != t.meta[i].y) return 6;
  }
  for (int a = 0; a < 3; a++)
    if (lb.lo[a] != t.leaf_box.lo[a] || lb.hi[a] != t.leaf_box.hi[a]) return 7;
  for (int d = 0; d < 24; d++)
    for (int a = 0; a < 3; a++)
      if (db[d].lo[a] != t.depth_box[d].lo[a] || db[d].hi[a] != t.depth_box[d].hi[a]) return 8;
  return 0;
}

// Incremental: transcode `old_raw` whole, then push `nranges` byte ranges [ranges[2k], ranges[2k+1]) of `new_raw` through
// gpu_diff_apply + gpu_patch, and compare the patched tree with a whole transcode of `new_raw`.
// out: [0] 0 = equivalent, [1..3] dirty / roots / appended summed over the ranges, [4] ranges that fell back, [5] reachable
// descriptors, [6] descriptors stored.
int emu_patch_check(const uint8_t *old_raw, uint64_t old_n, const uint8_t *new_raw, uint64_t new_n, const uint64_t *ranges, int nranges,
                    uint64_t out[8], int nthreads) {
  for (int k = 0; k < 8; k++) out[k] = 0;
  Transcoded t0, t1;
  std::string err;
  if (!transcode_stream(old_raw, old_n, t0, err, 2)) return 1;
  if (!transcode_stream(new_raw, new_n, t1, err, 2)) return 2;
  const uint64_t cap = t0.desc.size() + t0.desc.size() / 4 + 4096;
  std::vector<uint2> desc(cap), meta(cap);
  std::vector<uint32_t> ref(cap);
  std::vector<uint8_t> flag(cap, 0);
  std::copy(t0.desc.begin(), t0.desc.end(), desc.begin());
  std::copy(t0.refbase.begin(), t0.refbase.end(), ref.begin());
  std::copy(t0.meta.begin(), t0.meta.end(), meta.begin());
  uint64_t nd = t0.desc.size();
  CellBox lb = t0.leaf_box, db[24];
  for (int d = 0; d < 24; d++) db[d] = t0.depth_box[d];
  std::vector<uint8_t> raw(std::max(old_n, new_n) + 64, 0);
  memcpy(raw.data(), old_raw, old_n);
  uint64_t nbytes = old_n;
  simt::g_os_threads = nthreads;
  std::vector<uint8_t> arena(1 << 20);  // odd ranges use the arena (and spill from it), even ones plain allocations
  bool whole = false;
  for (int r = 0; r < nranges; r++) {
    const uint64_t start = ranges[2 * r], end = ranges[2 * r + 1];
    if (start >= end || end > new_n) return 3;
    const uint64_t before = nbytes;
    if (end > nbytes) nbytes = end;
    std::vector<uint8_t> bitmap((end - start + 7) / 8 + 16, 0);
    uint64_t span[2];
    if (gpu_diff_apply(raw.data(), new_raw + start, start, end, before, bitmap.data(), span, arena.data(), arena.size(), nullptr) != cudaSuccess) return 4;
    bool fallback = false;
    uint64_t stats[3];
    if (gpu_patch(raw.data(), nbytes, bitmap.data(), start, end, span, desc.data(), ref.data(), meta.data(), flag.data(), cap, &nd, &lb, db, &fallback,
                  stats, r % 2 ? arena.data() : nullptr, r % 2 ? arena.size() : 0, nullptr) != cudaSuccess)
      return 5;
    out[1] += stats[0]; out[2] += stats[1]; out[3] += stats[2];
    if (fallback) { out[4]++; whole = true; break; }
  }
  for (uint64_t i = 0; i < cap; i++)
    if (flag[i]) return 6;  // the scratch flags must be left clean
  if (whole) return 0;      // the product runs the whole transcode then; nothing to compare
  if (nbytes != new_n || memcmp(raw.data(), new_raw, new_n) != 0) { out[0] = 10; return 0; }  // the ranges did not cover every change
  uint64_t ca[3], cb[3];
  if (!canonical(desc, ref, nd, ca)) { out[0] = 11; return 0; }
  if (!canonical(t1.desc, t1.refbase, t1.desc.size(), cb)) return 7;
  out[5] = ca[0];
  out[6] = nd;
  if (ca[0] != cb[0] || ca[1] != cb[1] || ca[2] != cb[2]) { out[0] = 12; return 0; }
  if (!box_contains(lb, t1.leaf_box)) { out[0] = 13; return 0; }
  for (int d = 0; d < 24; d++)
    if (!box_contains(db[d], t1.depth_box[d])) { out[0] = 14; return 0; }
  return 0;
}

}  // extern "C"
