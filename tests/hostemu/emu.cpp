// emu.cpp -- TEST INFRASTRUCTURE: the device source of the trace path compiled for the host.
//
// g++ compiles svo_raytracer_b200/csrc/svo_trace.cuh (through cuda_host_shim.h) unchanged, so the statements the
// sm_100a kernels execute -- Trav::setup/run/step, finish_hit, pixel_begin/after_cast/store -- run here on the
// CPU, driven with the launch geometry and template selection of svo_kernels.cu.  Uses:
//   * tests/test_hostemu.py compares it with the oracle on every plane (the build container has no GPU; the
//     `-m gpu` tests do the same through the C ABI on the B200);
//   * emu_simt_* replays a frame warp by warp in lockstep and counts which of the loop's paths each warp
//     iteration has to issue -- the divergence model behind DESIGN.md's instruction budget.
// Never linked into libsvo_b200.so: the product path has no CPU implementation.
#include "cuda_host_shim.h"

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../svo_raytracer_b200/csrc/svo_trace.cuh"
#include "../../svo_raytracer_b200/csrc/svo_transcode.h"
#include "emu_scene.h"

using namespace svo;

static SceneView view_of(const emu_scene *s, const FrameParams *f) { return emu_view_of(s, f); }

template <class F>
static void parallel_rows(int y0, int y1, int nthreads, F fn) {
  if (nthreads <= 1) {
    for (int y = y0; y < y1; y++) fn(y);
    return;
  }
  std::atomic<int> next(y0);
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; i++)
    th.emplace_back([&] {
      for (;;) {
        const int y = next.fetch_add(1);
        if (y >= y1) break;
        fn(y);
      }
    });
  for (auto &t : th) t.join();
}

// the persistent / wavefront kernels' way through the same code: Trav::step + pixel_finish_cast
template <bool AUX>
static void shade_pixel_stepwise(const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H, int x, int y) {
  Pixel P;
  if (pixel_begin(f, pl, W, H, x, y, P)) {
    bool more;
    do {
      uint2 stk[kMaxScale + 1];
      Trav<false> T;
      T.setup(sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, nullptr);
      int status;
      do status = T.step(sc, stk, nullptr);
      while (status == TRAV_CONTINUE);
      more = pixel_finish_cast(sc, f, P, T.export_hit(status));
    } while (more);
  }
  pixel_store<AUX>(sc, f, pl, W, P);
}

extern "C" {

emu_scene *emu_scene_create(const uint8_t *nodes, uint64_t nbytes, char *err, int errlen) {
  emu_scene *s = new emu_scene;
  s->raw.assign(nodes, nodes + nbytes);
  std::string e;
  if (!transcode_stream(s->raw.data(), nbytes, s->t, e, 0)) {
    if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.c_str());
    delete s;
    return nullptr;
  }
  return s;
}
void emu_scene_destroy(emu_scene *s) { delete s; }
uint64_t emu_scene_ndesc(const emu_scene *s) { return s->t.desc.size(); }

// Rows [y0, y1) of one frame.  path: 0 = k_render_tile (Trav::run), 1 = the step-wise state machine of the
// persistent / wavefront kernels, 2 = k_render_stats (counters[3] filled).  box/aux select the template instance
// exactly as launch_render does (aux wins over box).
int emu_render(const emu_scene *s, const FrameParams *f, int W, int H, int y0, int y1, int path, int box, int aux, const float *beam,
               uint8_t *rgba8, float *depth, uint32_t *hit_id, uint32_t *iter, float *primary_t, float *radiance, uint64_t *counters,
               int nthreads) {
  const SceneView sc = view_of(s, f);
  Planes pl;
  pl.rgba8 = (uchar4 *)rgba8;
  pl.depth = depth;
  pl.beam = beam;
  pl.hit_id = hit_id;
  pl.iter = iter;
  pl.primary_t = primary_t;
  pl.radiance = (float4 *)radiance;
  if ((aux || path == 2) && !(hit_id && iter && primary_t && radiance)) return 1;
  std::atomic<uint64_t> c0(0), c1(0), c2(0);
  parallel_rows(y0, y1, nthreads, [&](int y) {
    RayStats rs;
    rs.casts = rs.iters = rs.record_bytes = 0u;
    for (int x = 0; x < W; x++) {
      if (path == 2) shade_pixel<false, true, true>(sc, *f, pl, W, H, x, y, &rs);
      else if (path == 1) { if (aux) shade_pixel_stepwise<true>(sc, *f, pl, W, H, x, y); else shade_pixel_stepwise<false>(sc, *f, pl, W, H, x, y); }
      else if (aux) shade_pixel<false, true, false, false>(sc, *f, pl, W, H, x, y);
      else if (box) shade_pixel<false, false, false, true>(sc, *f, pl, W, H, x, y);
      else shade_pixel<false, false, false, false>(sc, *f, pl, W, H, x, y);
    }
    c0 += rs.casts; c1 += rs.iters; c2 += rs.record_bytes;
  });
  if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; }
  return 0;
}

struct EmuRay { float o[3], d[3]; };
struct EmuHit { uint32_t id; float t; uint32_t value, iter; };

// k_cast_stream
int emu_cast(const emu_scene *s, const EmuRay *rays, uint64_t n, EmuHit *out, int maxDepth, int nthreads) {
  const SceneView sc = view_of(s, nullptr);
  const int chunks = (int)((n + 4095) / 4096);
  parallel_rows(0, chunks, nthreads, [&](int c) {
    const uint64_t a = (uint64_t)c * 4096, b = std::min<uint64_t>(n, a + 4096);
    for (uint64_t r = a; r < b; r++) {
      CastRes res;
      cast_res_clear(res);
      uint32_t loops = 0;
      const bool hit = cast_ray<false>(sc, mk3(rays[r].o[0], rays[r].o[1], rays[r].o[2]), mk3(rays[r].d[0], rays[r].d[1], rays[r].d[2]),
                                       maxDepth, false, 11, res, loops);
      out[r].id = hit ? res.pointer : kNoHit;
      out[r].t = hit ? res.t : 0.0f;
      out[r].value = hit ? res.value : 0u;
      out[r].iter = loops;
    }
  });
  return 0;
}

// k_beam
int emu_beam(const emu_scene *s, const FrameParams *fp, float *beam, int W, int H) {
  const FrameParams &f = *fp;
  const SceneView sc = view_of(s, nullptr);
  const int bw = W >> 2, bh = H >> 2;
  for (int gy = 0; gy < bh; gy++)
    for (int gx = 0; gx < bw; gx++) {
      const float fx = fdiv(fadd((float)(gx * 4), 0.5f), (float)W);
      const float fy = fdiv(fadd((float)(gy * 4), 0.5f), (float)H);
      vec3 dir;
      dir.x = mixf(mixf(f.l1[0], f.l2[0], fy), mixf(f.r1[0], f.r2[0], fy), fx);
      dir.y = mixf(mixf(f.l1[1], f.l2[1], fy), mixf(f.r1[1], f.r2[1], fy), fx);
      dir.z = mixf(mixf(f.l1[2], f.l2[2], fy), mixf(f.r1[2], f.r2[2], fy), fx);
      CastRes res;
      cast_res_clear(res);
      uint32_t loops = 0;
      const bool hit = cast_ray<false>(sc, mk3(f.camPos[0], f.camPos[1], f.camPos[2]), dir, f.maxDepth, false, f.coneDepth, res, loops);
      beam[(size_t)gy * (size_t)bw + (size_t)gx] = hit ? res.t : 0.0f;
    }
  return 0;
}

float emu_math(int fn, float x, float y) {
  switch (fn) {
    case 0: return det_sin(x);
    case 1: return det_cos(x);
    case 2: return det_acos(x);
    case 3: return det_exp(x);
    default: return det_rand(x, y);
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// SIMT divergence model.  A warp of the tile kernel owns an 8x4 pixel tile; its 32 lanes run the traversal loop in
// lockstep and the hardware re-converges them at the end of every iteration (BSSY/BSYNC around the loop body in
// the SASS), so a warp iteration issues the loop head once plus every path -- PUSH, ADVANCE, POP -- that at
// least one lane takes.  The sequence of paths a ray takes does not depend on the schedule, so it is recorded
// once per ray (from the device source, Trav::step) and replayed under different loop organisations with the
// per-path instruction counts read off the SASS.  Output: issue slots per organisation, to rank designs offline
// and to compare with ncu's smsp__inst_executed of the real kernel.
// ---------------------------------------------------------------------------------------------------------------
namespace {

enum Op : uint8_t { OP_PUSH = 0, OP_ADV = 1, OP_POP = 2, OP_EXIT_HIT = 3, OP_EXIT_MISS = 4 };

struct LaneCast {
  std::vector<uint8_t> ops;  // one entry per loop iteration (the last one is an exit)
  bool early = false;        // ended before the loop (outside the content box / NaN ray)
  float dy = 0.0f;           // direction of the ray (regrouping keys of the what-if model)
  uint8_t oct = 0;
};

// costs[]: 0 head, 1 push, 2 advance, 3 pop, 4 loop tail (re-convergence + back branch), 5 exit on the head/push side (hit),
//          6 exit on the pop side (miss), 7 per-cast code outside the loop (setup + finish + shading; issued once per warp cast),
//          8 one refill event of the dynamic-refill what-if, 9 its extra instructions per warp iteration (ballot + test)
struct Costs { double head, push, adv, pop, tail, exit_hit, exit_miss, outside, refill, refill_iter; };

struct Tally {
  double slots[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // per organisation
  double thread_ops = 0;                       // sum over lanes of their own path costs (what 100 % lane utilisation would issue x32)
  double longest_lane = 0;                     // per warp cast: the most expensive lane alone (trip-count divergence only)
  uint64_t iters = 0, casts = 0, warp_casts = 0, pushes = 0, advs = 0, pops = 0, early = 0;
  uint64_t warp_iters0 = 0, warp_iters_any[3] = {0, 0, 0};
  // what-if: the casts after the first (bounce / shadow rays) of a group of G neighbouring tiles re-dealt to warps
  // [g][0] as shipped, [g][1] sorted by true iteration count (bound), [g][2] by dir.y, [g][3] by (octant, dir.y), [g][4] by octant
  double regroup[3][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
  double refill[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};  // [g][variant]: dynamic refill, see kRefillVariants
};
constexpr int kGroupSizes[3] = {4, 16, 64};
// {workers as a fraction of the group's warps (1/x), idle-lane threshold}
constexpr int kRefillVariants[4][2] = {{2, 8}, {2, 16}, {4, 8}, {4, 16}};

inline double own_cost(const Costs &c, uint8_t op) {
  switch (op) {
    case OP_PUSH: return c.head + c.push + c.tail;
    case OP_ADV: return c.head + c.adv + c.tail;
    case OP_POP: return c.head + c.adv + c.pop + c.tail;
    case OP_EXIT_HIT: return c.head + c.exit_hit;
    default: return c.head + c.adv + c.exit_miss;
  }
}

// organisation 0: the shipped loop -- every lane does one iteration per warp iteration
double org_if_if(const Costs &c, const std::vector<const LaneCast *> &lanes, Tally &t) {
  size_t pos[32] = {0};
  double s = 0;
  for (;;) {
    bool any[5] = {false, false, false, false, false};
    bool alive = false;
    for (size_t l = 0; l < lanes.size(); l++)
      if (pos[l] < lanes[l]->ops.size()) { any[lanes[l]->ops[pos[l]++]] = true; alive = true; }
    if (!alive) break;
    t.warp_iters0++;
    if (any[OP_PUSH]) t.warp_iters_any[0]++;
    if (any[OP_ADV] || any[OP_POP]) t.warp_iters_any[1]++;
    if (any[OP_POP]) t.warp_iters_any[2]++;
    s += c.head + c.tail;
    if (any[OP_PUSH]) s += c.push;
    if (any[OP_EXIT_HIT]) s += c.exit_hit;
    if (any[OP_ADV] || any[OP_POP] || any[OP_EXIT_MISS]) s += c.adv;
    if (any[OP_POP]) s += c.pop;
    if (any[OP_EXIT_MISS]) s += c.exit_miss;
  }
  return s;
}

// organisation 1: while-while -- an inner loop of PUSHes (lanes that want to ADVANCE wait), then an inner loop of
// ADVANCE/POPs (lanes that want to PUSH wait); `max_push` / `max_adv` bound the inner trip counts (0 = unbounded)
double org_while_while(const Costs &c, const std::vector<const LaneCast *> &lanes, int max_push, int max_adv) {
  size_t pos[32] = {0};
  double s = 0;
  const double phase = 2;  // ballot/branch per inner-loop trip
  for (;;) {
    bool alive = false;
    for (size_t l = 0; l < lanes.size(); l++) alive |= pos[l] < lanes[l]->ops.size();
    if (!alive) break;
    for (int trip = 0; max_push == 0 || trip < max_push; trip++) {  // PUSH phase (hit exits are found on this side)
      bool any = false, any_exit = false;
      for (size_t l = 0; l < lanes.size(); l++)
        if (pos[l] < lanes[l]->ops.size()) {
          const uint8_t op = lanes[l]->ops[pos[l]];
          if (op == OP_PUSH) { any = true; pos[l]++; }
          else if (op == OP_EXIT_HIT) { any_exit = true; pos[l]++; }
        }
      if (!any && !any_exit) { s += trip == 0 ? c.head + phase : 0; break; }
      s += c.head + phase + (any ? c.push : 0) + (any_exit ? c.exit_hit : 0);
    }
    for (int trip = 0; max_adv == 0 || trip < max_adv; trip++) {  // ADVANCE / POP phase
      bool any = false, any_pop = false, any_exit = false;
      for (size_t l = 0; l < lanes.size(); l++)
        if (pos[l] < lanes[l]->ops.size()) {
          const uint8_t op = lanes[l]->ops[pos[l]];
          if (op == OP_ADV) { any = true; pos[l]++; }
          else if (op == OP_POP) { any = any_pop = true; pos[l]++; }
          else if (op == OP_EXIT_MISS) { any = any_exit = true; pos[l]++; }
        }
      if (!any) { s += trip == 0 ? c.head + phase : 0; break; }
      s += c.head + phase + c.adv + (any_pop ? c.pop : 0) + (any_exit ? c.exit_miss : 0);
    }
    s += c.tail;
  }
  return s;
}

// what-if organisation: `workers` warps share a queue of rays (the bounce casts of a group of tiles); a warp whose idle
// lanes reach `threshold` leaves the loop, takes rays from the queue (costs `refill` slots: loop exit/entry + Trav::setup
// for the new lanes; the finished lanes' end states are parked for their owners) and goes on.  Returns issue slots.
double org_refill(const Costs &c, const std::vector<const LaneCast *> &queue, int workers, int threshold) {
  const double refill = c.refill;
  struct Lane { const LaneCast *lc = nullptr; size_t pos = 0; };
  struct Worker { Lane lane[32]; double t = 0; bool done = false; };
  std::vector<Worker> w((size_t)workers);
  size_t head = 0;
  double total = 0;
  auto fill = [&](Worker &k) {
    int taken = 0;
    for (int l = 0; l < 32 && head < queue.size(); l++)
      if (!k.lane[l].lc) { k.lane[l].lc = queue[head++]; k.lane[l].pos = 0; taken++; }
    return taken;
  };
  for (auto &k : w) if (fill(k)) { k.t += refill; total += refill; } else k.done = true;
  for (;;) {
    Worker *k = nullptr;
    for (auto &x : w) if (!x.done && (!k || x.t < k->t)) k = &x;
    if (!k) break;
    bool any[5] = {false, false, false, false, false};
    int active = 0;
    for (int l = 0; l < 32; l++) {
      Lane &ln = k->lane[l];
      if (!ln.lc) continue;
      if (ln.pos < ln.lc->ops.size()) { any[ln.lc->ops[ln.pos++]] = true; active++; }
      if (ln.pos >= ln.lc->ops.size()) ln.lc = nullptr;
    }
    double s = 0;
    if (active) {
      s = c.head + c.tail + c.refill_iter;
      if (any[OP_PUSH]) s += c.push;
      if (any[OP_EXIT_HIT]) s += c.exit_hit;
      if (any[OP_ADV] || any[OP_POP] || any[OP_EXIT_MISS]) s += c.adv;
      if (any[OP_POP]) s += c.pop;
      if (any[OP_EXIT_MISS]) s += c.exit_miss;
    }
    int idle = 0, busy = 0;
    for (int l = 0; l < 32; l++) (k->lane[l].lc ? busy : idle)++;
    if ((idle >= threshold || busy == 0) && head < queue.size()) { fill(*k); s += refill; }
    else if (busy == 0) k->done = true;
    k->t += s;
    total += s;
  }
  return total;
}

}  // namespace

extern "C" {

// Replays rows [y0, y1) of a frame warp by warp (8x4 pixel tiles, the tile kernel's mapping).  out[0..7] = issue slots
// of the loop organisations (0 shipped if-if, 1 while-while unbounded, 2 while-while 1 push trip / 1 advance trip
// (alternating), 3 while-while unbounded pushes / 1 advance, 4 while-while 1 push / unbounded advances, 5 bounded
// 4/2), out[8] = thread ops / 32 (perfect lane utilisation), out[9] = longest lane per warp cast (trip-count
// divergence only), out[10..] = counts: casts, iterations, pushes, advances, pops, warp casts, warp iterations of
// the shipped loop, and of those the ones that issue PUSH / ADVANCE / POP.
int emu_simt(const emu_scene *s, const FrameParams *fp, int W, int H, int y0, int y1, int box, const double *costs, double *out,
             int nthreads, int tile_w) {
  if (tile_w != 4 && tile_w != 8 && tile_w != 16 && tile_w != 32) tile_w = 8;  // warp = tile_w x (32 / tile_w) pixels; the kernel uses 8x4
  const int tile_h = 32 / tile_w, tw_shift = tile_w == 4 ? 2 : tile_w == 8 ? 3 : tile_w == 16 ? 4 : 5;
  const FrameParams &f = *fp;
  const SceneView sc = view_of(s, fp);
  Costs c{costs[0], costs[1], costs[2], costs[3], costs[4], costs[5], costs[6], costs[7], costs[8], costs[9]};
  const int tiles_x = (W + tile_w - 1) / tile_w, tiles_y0 = y0 / tile_h, tiles_y1 = (y1 + tile_h - 1) / tile_h;
  std::vector<Tally> tallies((size_t)std::max(1, nthreads));
  std::atomic<int> next(tiles_y0);
  std::vector<uint8_t> rgba((size_t)W * H * 4);
  std::vector<float> depth((size_t)W * H);
  Planes pl;
  pl.rgba8 = (uchar4 *)rgba.data();
  pl.depth = depth.data();
  pl.beam = nullptr;
  pl.hit_id = nullptr; pl.iter = nullptr; pl.primary_t = nullptr; pl.radiance = nullptr;
  auto work = [&](int tid) {
    Tally &t = tallies[(size_t)tid];
    for (;;) {
      const int ty = next.fetch_add(1);
      if (ty >= tiles_y1) break;
      std::vector<std::vector<LaneCast>> pool[3];  // per group size: [cast index - 1] -> rays of the group so far
      int pooled[3] = {0, 0, 0};
      auto flush = [&](int g) {
        for (auto &rays : pool[g]) {
          if (rays.empty()) continue;
          Tally scratch;
          for (int order = 0; order < 5; order++) {
            std::vector<const LaneCast *> v;
            for (const LaneCast &lc : rays) v.push_back(&lc);
            if (order == 1) std::stable_sort(v.begin(), v.end(), [](const LaneCast *a, const LaneCast *b) { return a->ops.size() < b->ops.size(); });
            if (order == 2) std::stable_sort(v.begin(), v.end(), [](const LaneCast *a, const LaneCast *b) { return a->dy < b->dy; });
            if (order == 3) std::stable_sort(v.begin(), v.end(), [](const LaneCast *a, const LaneCast *b) { return a->oct != b->oct ? a->oct < b->oct : a->dy < b->dy; });
            if (order == 4) std::stable_sort(v.begin(), v.end(), [](const LaneCast *a, const LaneCast *b) { return a->oct < b->oct; });
            for (size_t i = 0; i < v.size(); i += 32) {
              std::vector<const LaneCast *> w(v.begin() + (long)i, v.begin() + (long)std::min(v.size(), i + 32));
              t.regroup[g][order] += org_if_if(c, w, scratch) + c.outside;
            }
          }
        }
        for (auto &rays : pool[g]) {
          if (rays.empty()) continue;
          std::vector<const LaneCast *> v;
          for (const LaneCast &lc : rays) v.push_back(&lc);
          for (int r = 0; r < 4; r++) {
            const int warps = (int)((v.size() + 31) / 32);
            const int workers = std::max(1, warps / kRefillVariants[r][0]);
            // code outside the loop: the shipped kernel pays c.outside per warp cast; here the shading still runs once per 32 rays
            t.refill[g][r] += org_refill(c, v, workers, kRefillVariants[r][1]) + c.outside * warps;
          }
        }
        pool[g].clear();
        pooled[g] = 0;
      };
      for (int tx = 0; tx < tiles_x; tx++) {
        std::vector<std::vector<LaneCast>> lane_casts(32);  // [lane][cast]
        for (int l = 0; l < 32; l++) {
          const int x = tx * tile_w + (l & (tile_w - 1)), y = ty * tile_h + (l >> tw_shift);
          if (x >= W || y < y0 || y >= y1) continue;
          Pixel P;
          if (!pixel_begin(f, pl, W, H, x, y, P)) continue;
          bool more;
          do {
            lane_casts[l].emplace_back();
            LaneCast &lc = lane_casts[l].back();
            lc.dy = P.dir.y;
            lc.oct = (uint8_t)((P.dir.x > 0.0f ? 1 : 0) | (P.dir.y > 0.0f ? 2 : 0) | (P.dir.z > 0.0f ? 4 : 0));
            uint2 stk[kMaxScale + 1];
            int status;
            if (box) {
              Trav<false, false, true> T;
              T.setup(sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, nullptr);
              if (T.outside_box() || T.nan_ray(nullptr)) { status = TRAV_MISS; lc.early = true; }
              else for (;;) {
                const int s0 = T.scale;
                status = T.step(sc, stk, nullptr);
                if (status != TRAV_CONTINUE) { lc.ops.push_back(status == TRAV_HIT ? OP_EXIT_HIT : OP_EXIT_MISS); break; }
                lc.ops.push_back(T.scale < s0 ? OP_PUSH : (T.scale > s0 ? OP_POP : OP_ADV));
              }
              more = pixel_finish_cast(sc, f, P, T.export_hit(status));
            } else {
              Trav<false> T;
              T.setup(sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, nullptr);
              if (T.nan_ray(nullptr)) { status = TRAV_MISS; lc.early = true; }
              else for (;;) {
                const int s0 = T.scale;
                status = T.step(sc, stk, nullptr);
                if (status != TRAV_CONTINUE) { lc.ops.push_back(status == TRAV_HIT ? OP_EXIT_HIT : OP_EXIT_MISS); break; }
                lc.ops.push_back(T.scale < s0 ? OP_PUSH : (T.scale > s0 ? OP_POP : OP_ADV));
              }
              more = pixel_finish_cast(sc, f, P, T.export_hit(status));
            }
          } while (more);
        }
        for (size_t k = 0;; k++) {
          std::vector<const LaneCast *> lanes;
          for (int l = 0; l < 32; l++)
            if (lane_casts[l].size() > k) lanes.push_back(&lane_casts[l][k]);
          if (lanes.empty()) break;
          t.warp_casts++;
          double longest = 0;
          for (const LaneCast *lc : lanes) {
            double own = 0;
            for (uint8_t op : lc->ops) {
              own += own_cost(c, op);
              t.pushes += op == OP_PUSH; t.advs += op == OP_ADV; t.pops += op == OP_POP;
            }
            t.iters += lc->ops.size();
            t.casts++;
            t.early += lc->early;
            t.thread_ops += own + c.outside;
            longest = std::max(longest, own);
          }
          t.longest_lane += longest + c.outside;
          t.slots[0] += org_if_if(c, lanes, t) + c.outside;
          t.slots[1] += org_while_while(c, lanes, 0, 0) + c.outside;
          t.slots[2] += org_while_while(c, lanes, 1, 1) + c.outside;
          t.slots[3] += org_while_while(c, lanes, 0, 1) + c.outside;
          t.slots[4] += org_while_while(c, lanes, 1, 0) + c.outside;
          t.slots[5] += org_while_while(c, lanes, 4, 2) + c.outside;
        }
        for (int g = 0; g < 3; g++) {
          for (int l = 0; l < 32; l++)
            for (size_t k = 1; k < lane_casts[l].size(); k++) {
              if (pool[g].size() < k) pool[g].resize(k);
              pool[g][k - 1].push_back(lane_casts[l][k]);
            }
          if (++pooled[g] == kGroupSizes[g] || tx + 1 == tiles_x) flush(g);
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(work, i);
  work(0);
  for (auto &x : th) x.join();
  Tally sum;
  for (const Tally &t : tallies) {
    for (int i = 0; i < 8; i++) sum.slots[i] += t.slots[i];
    sum.thread_ops += t.thread_ops; sum.longest_lane += t.longest_lane;
    sum.iters += t.iters; sum.casts += t.casts; sum.warp_casts += t.warp_casts;
    sum.pushes += t.pushes; sum.advs += t.advs; sum.pops += t.pops; sum.early += t.early;
    for (int g = 0; g < 3; g++) for (int o = 0; o < 5; o++) sum.regroup[g][o] += t.regroup[g][o];
    for (int g = 0; g < 3; g++) for (int o = 0; o < 4; o++) sum.refill[g][o] += t.refill[g][o];
    sum.warp_iters0 += t.warp_iters0;
    for (int i = 0; i < 3; i++) sum.warp_iters_any[i] += t.warp_iters_any[i];
  }
  for (int i = 0; i < 8; i++) out[i] = sum.slots[i];
  out[8] = sum.thread_ops / 32.0;
  out[9] = sum.longest_lane;
  out[10] = (double)sum.casts; out[11] = (double)sum.iters; out[12] = (double)sum.pushes; out[13] = (double)sum.advs;
  out[14] = (double)sum.pops; out[15] = (double)sum.warp_casts; out[16] = (double)sum.warp_iters0;
  out[17] = (double)sum.warp_iters_any[0]; out[18] = (double)sum.warp_iters_any[1]; out[19] = (double)sum.warp_iters_any[2];
  for (int g = 0; g < 3; g++) for (int o = 0; o < 5; o++) out[20 + g * 5 + o] = sum.regroup[g][o];
  for (int g = 0; g < 3; g++) for (int o = 0; o < 4; o++) out[35 + g * 4 + o] = sum.refill[g][o];
  out[47] = (double)sum.early;  // casts that ended before the loop (outside the content box, NaN rays)
  return 0;
}

}  // extern "C"

// Path string of every cast of pixel (x, y): 'P' PUSH, 'A' ADVANCE, 'Q' ADVANCE+POP, 'H' exit hit, 'M' exit miss, '|' between casts;
// each op is followed by the scale digit-letter ('a' + scale) it ran at.  For eyeballing what rays do (tools/simt_model.py --pixel).
extern "C" int emu_pixel_ops(const emu_scene *s, const FrameParams *fp, int W, int H, int x, int y, int box, char *buf, int cap) {
  const FrameParams &f = *fp;
  const SceneView sc = view_of(s, fp);
  std::vector<uint8_t> rgba((size_t)W * H * 4);
  std::vector<float> depth((size_t)W * H);
  Planes pl;
  pl.rgba8 = (uchar4 *)rgba.data(); pl.depth = depth.data(); pl.beam = nullptr;
  pl.hit_id = nullptr; pl.iter = nullptr; pl.primary_t = nullptr; pl.radiance = nullptr;
  std::string out;
  Pixel P;
  if (pixel_begin(f, pl, W, H, x, y, P)) {
    bool more;
    do {
      uint2 stk[kMaxScale + 1];
      int status;
      Trav<false, false, true> T;
      T.setup(sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, nullptr);
      if ((box && T.outside_box()) || T.nan_ray(nullptr)) { status = TRAV_MISS; out += "X"; }
      else for (;;) {
        if (!box) T.tb_out = __uint_as_float(0x7f800000u);
        const int s0 = T.scale;
        status = T.step(sc, stk, nullptr);
        if (status != TRAV_CONTINUE) { out += status == TRAV_HIT ? 'H' : 'M'; out += (char)('a' + s0); break; }
        out += T.scale < s0 ? 'P' : (T.scale > s0 ? 'Q' : 'A');
        out += (char)('a' + s0);
      }
      more = pixel_finish_cast(sc, f, P, T.export_hit(status));
      out += '|';
    } while (more);
  }
  snprintf(buf, (size_t)cap, "%s", out.c_str());
  return (int)out.size();
}

// The divergence model for a ray stream (svo_cast): rays in the given order, 32 consecutive rays per warp.
// out[0] = issue slots of the grid-stride kernel (one ray per lane per pass, the warp waits for its slowest ray),
// out[1] = of the persistent kernel (k_cast_stream_persistent: chunks of 256 rays per warp, lanes re-armed when >= 8 are
// free), out[2] = 32 busy lanes, out[3] = casts, out[4] = iterations, out[5] = slowest lane per warp pass.
extern "C" int emu_simt_stream(const emu_scene *s, const EmuRay *rays, const uint32_t *order, uint64_t n, int maxDepth, const double *costs,
                               double *out, int nthreads) {
  const SceneView sc = view_of(s, nullptr);
  Costs c{costs[0], costs[1], costs[2], costs[3], costs[4], costs[5], costs[6], costs[7], costs[8], costs[9]};
  const uint64_t chunk = 256;
  const uint64_t nchunks = (n + chunk - 1) / chunk;
  std::vector<Tally> tallies((size_t)std::max(1, nthreads));
  std::vector<double> refill((size_t)std::max(1, nthreads), 0.0);
  std::atomic<uint64_t> next(0);
  auto work = [&](int tid) {
    Tally &t = tallies[(size_t)tid];
    for (;;) {
      const uint64_t ci = next.fetch_add(1);
      if (ci >= nchunks) break;
      const uint64_t a = ci * chunk, b = std::min(n, a + chunk);
      std::vector<LaneCast> lc((size_t)(b - a));
      for (uint64_t i = a; i < b; i++) {
        const EmuRay &r = rays[order ? order[i] : i];
        LaneCast &l = lc[(size_t)(i - a)];
        uint2 stk[kMaxScale + 1];
        Trav<false> T;
        T.setup(sc, mk3(r.o[0], r.o[1], r.o[2]), mk3(r.d[0], r.d[1], r.d[2]), maxDepth, false, 11, nullptr);
        if (T.nan_ray(nullptr)) { l.early = true; continue; }
        for (;;) {
          const int s0 = T.scale;
          const int status = T.step(sc, stk, nullptr);
          if (status != TRAV_CONTINUE) { l.ops.push_back(status == TRAV_HIT ? OP_EXIT_HIT : OP_EXIT_MISS); break; }
          l.ops.push_back(T.scale < s0 ? OP_PUSH : (T.scale > s0 ? OP_POP : OP_ADV));
        }
      }
      std::vector<const LaneCast *> all;
      for (const LaneCast &l : lc) {
        all.push_back(&l);
        double own = 0;
        for (uint8_t op : l.ops) own += own_cost(c, op);
        t.thread_ops += own + c.outside;
        t.iters += l.ops.size();
        t.casts++;
      }
      for (size_t i = 0; i < all.size(); i += 32) {
        std::vector<const LaneCast *> w(all.begin() + (long)i, all.begin() + (long)std::min(all.size(), i + 32));
        t.slots[0] += org_if_if(c, w, t) + c.outside;
        double longest = 0;
        for (const LaneCast *l : w) { double own = 0; for (uint8_t op : l->ops) own += own_cost(c, op); longest = std::max(longest, own); }
        t.longest_lane += longest + c.outside;
      }
      // the persistent kernel: one warp works through the chunk; the hit records are written by the finishing lanes
      refill[(size_t)tid] += org_refill(c, all, 1, 8) + c.outside * (double)((all.size() + 31) / 32) * 0.25;
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(work, i);
  work(0);
  for (auto &x : th) x.join();
  for (int i = 0; i < 6; i++) out[i] = 0;
  for (size_t i = 0; i < tallies.size(); i++) {
    out[0] += tallies[i].slots[0];
    out[1] += refill[i];
    out[2] += tallies[i].thread_ops / 32.0;
    out[3] += (double)tallies[i].casts;
    out[4] += (double)tallies[i].iters;
    out[5] += tallies[i].longest_lane;
  }
  return 0;
}
