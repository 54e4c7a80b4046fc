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

using namespace svo;

struct emu_scene {
  std::vector<uint8_t> raw;
  Transcoded t;
};

static SceneView view_of(const emu_scene *s, const FrameParams *f) {
  SceneView v;
  content_box(s->t.leaf_box, s->t.depth_box, f ? f->maxDepth : -1, f ? f->coneDepth : -1, v.box_lo, v.box_hi);
  v.desc = s->t.desc.data();
  v.refbase = s->t.refbase.data();
  v.raw = s->raw.data();
  v.nbytes = s->raw.size();
  v.ndesc = (uint32_t)s->t.desc.size();
  uint32_t w0 = 0;
  memcpy(&w0, s->raw.data(), std::min<size_t>(4, s->raw.size()));
  v.first_word_zero = w0 == 0u;
  v.top = nullptr;
  v.ntop = 0;
  return v;
}

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
