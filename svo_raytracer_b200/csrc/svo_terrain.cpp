// svo_terrain.cpp -- deterministic synthetic heightmap + material map
// (svo_terrain_generate in include/svo_b200.h).
//
// Stand-in for the reference's input images: assets/heightmaps/*.png (16-bit,
// loaded at src/engine/Octree.java:206-215) and assets/matmaps/**/materials.png
// (8-bit ids 1 stone / 2 scree / 3 grass, Octree.java:217-226).  The 8192^2
// maps the reference names are absent upstream, so worlds of any size are
// generated here: ridged value-noise fBm from an integer lattice hash, IEEE
// double arithmetic only (no libm), hence bit-reproducible for (n, seed) on any
// host and for any thread count.
#include "../../include/svo_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace {

inline double lattice(uint32_t ix, uint32_t iy, uint32_t salt) {
  uint32_t h = (ix * 0x9E3779B1u) ^ (iy * 0x85EBCA77u);
  h ^= salt;
  h ^= h >> 15; h *= 0x2C1B3C6Du;
  h ^= h >> 12; h *= 0x297A2D39u;
  h ^= h >> 15;
  return (double)h * (1.0 / 4294967296.0);
}

struct Octave { int cells; uint32_t salt; double amp; bool ridged; };

template <class F>
void parallel_rows(int n, int nthreads, F f) {
  std::vector<std::thread> pool;
  auto work = [&](int t) { for (int z = t; z < n; z += nthreads) f(z); };
  for (int t = 1; t < nthreads; t++) pool.emplace_back(work, t);
  work(0);
  for (auto &th : pool) th.join();
}

}  // namespace

extern "C" int svo_terrain_generate(int n, int seed, uint16_t *height, uint8_t *mat, int nthreads) {
  if (n < 4 || (n & (n - 1)) || !height || !mat) return SVO_ERR_INVALID;
  if (nthreads < 1) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n;

  std::vector<Octave> octs;
  double amp = 1.0, total = 0.0;
  for (int cells = 4, o = 0; cells <= n / 2 && o < 11; cells *= 2, o++) {
    octs.push_back({cells, (uint32_t)(seed * 7919 + o * 104729), amp, o >= 2});
    total += amp;
    amp *= 0.5;
  }
  const size_t nn = (size_t)n * (size_t)n;
  std::vector<double> acc(nn);
  // per-axis interpolation tables: cell index and smoothstep weight of every texel centre
  std::vector<std::vector<int>> cell(octs.size(), std::vector<int>(n));
  std::vector<std::vector<double>> wgt(octs.size(), std::vector<double>(n));
  for (size_t o = 0; o < octs.size(); o++)
    for (int i = 0; i < n; i++) {
      const double t = ((double)i + 0.5) * ((double)octs[o].cells / (double)n);
      const double fl = std::floor(t);
      const double f = t - fl;
      cell[o][i] = (int)fl;
      wgt[o][i] = f * f * (3.0 - 2.0 * f);
    }
  std::vector<double> row_lo(n, 1e300), row_hi(n, -1e300);
  parallel_rows(n, nthreads, [&](int z) {
    double lo = 1e300, hi = -1e300;
    for (int x = 0; x < n; x++) {
      double s = 0.0;
      for (size_t o = 0; o < octs.size(); o++) {
        const uint32_t cx = (uint32_t)cell[o][x], cz = (uint32_t)cell[o][z];
        const double wx = wgt[o][x], wz = wgt[o][z];
        const double v00 = lattice(cx, cz, octs[o].salt), v10 = lattice(cx + 1, cz, octs[o].salt);
        const double v01 = lattice(cx, cz + 1, octs[o].salt), v11 = lattice(cx + 1, cz + 1, octs[o].salt);
        const double top = v00 + (v10 - v00) * wx, bot = v01 + (v11 - v01) * wx;
        double v = top + (bot - top) * wz;
        if (octs[o].ridged) v = 1.0 - std::fabs(2.0 * v - 1.0);
        s += octs[o].amp * v;
      }
      s /= total;
      acc[(size_t)z * n + x] = s;
      lo = std::min(lo, s);
      hi = std::max(hi, s);
    }
    row_lo[z] = lo;
    row_hi[z] = hi;
  });
  double lo = 1e300, hi = -1e300;
  for (int z = 0; z < n; z++) { lo = std::min(lo, row_lo[z]); hi = std::max(hi, row_hi[z]); }
  const double inv = 1.0 / (hi - lo);
  parallel_rows(n, nthreads, [&](int z) {
    for (int x = 0; x < n; x++) {
      double a = (acc[(size_t)z * n + x] - lo) * inv;
      a = a * (0.35 + 0.65 * a);  // flatten valleys, keep sharp peaks
      // span of assets/heightmaps/nz.png: 1957 .. 58795
      height[(size_t)z * n + x] = (uint16_t)std::floor(1957.0 + a * (58795.0 - 1957.0));
    }
  });
  // materials from height and slope (wrap-around neighbours), resolution independent classes
  parallel_rows(n, nthreads, [&](int z) {
    const int zm = (z + n - 1) & (n - 1), zp = (z + 1) & (n - 1);
    for (int x = 0; x < n; x++) {
      const int xm = (x + n - 1) & (n - 1), xp = (x + 1) & (n - 1);
      const int h = height[(size_t)z * n + x];
      const int gx = std::abs((int)height[(size_t)z * n + xp] - (int)height[(size_t)z * n + xm]);
      const int gz = std::abs((int)height[(size_t)zp * n + x] - (int)height[(size_t)zm * n + x]);
      const long long slope = (long long)(gx + gz) * n / 1024;
      uint8_t m = 3;
      if (slope > 500 || h > 30000) m = 2;
      if (slope > 1100 || h > 44000) m = 1;
      mat[(size_t)z * n + x] = m;
    }
  });
  return SVO_OK;
}
