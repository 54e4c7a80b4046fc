/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE.  Dispatch loop and extern "C" entry points around the reference's own
 * shaders compiled for the CPU (oracle/build_ref.py generates ref_svotrace_gen.inc / ref_svobeam_gen.inc from
 * /root/reference/src/shaders/svotrace.comp and svobeam.comp at build time; see that script for the rewrite).
 *
 * What this file does is what the GL host does around the shader: set the uniforms (Main.java:259-283), bind the
 * images (Main.java:62-86) and the node buffer (Main.java:122), and run main() once per invocation
 * (Renderer.dispatchCompute, Main.java:265,285).  It contains no traversal or shading code.
 *
 * Entry points mirror oracle/svo_oracle.h so that tests can run both on the same inputs.  Extra planes (hit_id,
 * iter, primary_t) come from the cast log build_ref.py's wrapper keeps (the shader's own pointer/iter store is
 * commented out, svotrace.comp:728), with the oracle's U7 convention.
 */
#include <pthread.h>

#include <cstdlib>
#include <mutex>

#include "glsl_shim.h"
#include "svo_oracle.h"

#define REF_MAX_LOG 8
struct ref_cast_log {
  bool hit;
  uint32_t loop_iter, pointer;
  float t;
  uint32_t value, iter;
};
struct ref_invocation_base {
  glsl::uvec3_id gl_GlobalInvocationID;
  uint32_t probe_iter = 0;
  int n_casts = 0;
  ref_cast_log cast_log[REF_MAX_LOG] = {};
};

#include "ref_svotrace_gen.inc"
#include "ref_svobeam_gen.inc"
namespace ref_svotrace = glsl::ref_svotrace; /* generated inside namespace glsl so that sin/abs/min... resolve to the shim */
namespace ref_svobeam = glsl::ref_svobeam;

namespace {

std::mutex g_lock; /* the uniforms are namespace-scope variables: one dispatch at a time */

template <class CR>
void to_shader(const svo_o_cast_result &s, CR &r) {
  r.value = s.value; r.pointer = s.pointer; r.iter = s.iter; r.t = s.t;
  r.hitPos = glsl::vec3(s.hitPos[0], s.hitPos[1], s.hitPos[2]);
  r.scale = s.scale;
  r.debugColor = glsl::vec3(s.debugColor[0], s.debugColor[1], s.debugColor[2]);
  r.normal = glsl::vec3(s.normal[0], s.normal[1], s.normal[2]);
  r.voxelPos = glsl::vec3(s.voxelPos[0], s.voxelPos[1], s.voxelPos[2]);
}
template <class CR>
void from_shader(const CR &r, svo_o_cast_result &s) {
  s.value = r.value; s.pointer = r.pointer; s.iter = r.iter; s.t = r.t;
  s.hitPos[0] = r.hitPos.x; s.hitPos[1] = r.hitPos.y; s.hitPos[2] = r.hitPos.z;
  s.scale = r.scale;
  s.debugColor[0] = r.debugColor.x; s.debugColor[1] = r.debugColor.y; s.debugColor[2] = r.debugColor.z;
  s.normal[0] = r.normal.x; s.normal[1] = r.normal.y; s.normal[2] = r.normal.z;
  s.voxelPos[0] = r.voxelPos.x; s.voxelPos[1] = r.voxelPos.y; s.voxelPos[2] = r.voxelPos.z;
}

glsl::vec3 v3(const float *p) { return glsl::vec3(p[0], p[1], p[2]); }

void bind_trace(const uint8_t *nodes, uint64_t nbytes, int maxDepth, int coneDepth, int casts) {
  ref_svotrace::octreeBuffer.bytes = nodes;
  ref_svotrace::octreeBuffer.nbytes = nbytes;
  ref_svotrace::bufferEnd = (int)nbytes; /* Main.java:281 */
  ref_svotrace::ref_max_depth = maxDepth;
  ref_svotrace::ref_cone_depth = coneDepth;
  ref_svotrace::ref_casts = casts;
}

struct job {
  int width, y0, y1, tid, nthreads;
  float *depth_plane; /* r32f image target */
  uint32_t *hit_id, *iter;
  float *primary_t;
};

void *render_worker(void *arg) {
  job *j = (job *)arg;
  for (int y = j->y0 + j->tid; y < j->y1; y += j->nthreads)
    for (int x = 0; x < j->width; x++) {
      ref_svotrace::Invocation inv; /* fresh globals per invocation (R2) */
      inv.gl_GlobalInvocationID.xy = glsl::uvec2((uint32_t)x, (uint32_t)y);
      inv.main();
      size_t p = (size_t)y * (size_t)j->width + (size_t)x;
      const ref_cast_log &c = inv.cast_log[0];
      bool any = inv.n_casts > 0;
      if (j->hit_id) j->hit_id[p] = (any && c.hit) ? c.pointer : SVO_O_NO_HIT;
      if (j->iter) j->iter[p] = any ? c.loop_iter : 0u;
      if (j->primary_t) j->primary_t[p] = (any && c.hit) ? c.t : 0.0f;
    }
  return nullptr;
}

struct rays_job {
  const svo_o_ray *rays;
  svo_o_hit *out;
  uint64_t n;
  int maxDepth, tid, nthreads;
};

void *rays_worker(void *arg) {
  rays_job *j = (rays_job *)arg;
  uint64_t per = (j->n + (uint64_t)j->nthreads - 1) / (uint64_t)j->nthreads;
  uint64_t lo = per * (uint64_t)j->tid, hi = lo + per;
  if (hi > j->n) hi = j->n;
  for (uint64_t i = lo; i < hi; i++) {
    ref_svotrace::Invocation inv;
    ref_svotrace::Invocation::castResult res = {};
    glsl::vec3 d = v3(j->rays[i].d);
    bool hit = inv.intersectOctree(v3(j->rays[i].o), d, 1.0f / d, res, j->maxDepth, false);
    if (hit) {
      j->out[i].id = res.pointer; j->out[i].t = res.t; j->out[i].value = res.value; j->out[i].iter = res.iter;
    } else {
      j->out[i].id = SVO_O_NO_HIT; j->out[i].t = 0.0f; j->out[i].value = 0; j->out[i].iter = inv.cast_log[0].loop_iter;
    }
  }
  return nullptr;
}

template <class J>
void run_threads(J *jobs, int nthreads, void *(*fn)(void *)) {
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
  if (nthreads == 1) fn(&jobs[0]);
  else {
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], nullptr, fn, &jobs[t]);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], nullptr);
  }
  free(th);
}

int clamp_threads(int n) { return n < 1 ? 1 : (n > 256 ? 256 : n); }

}  // namespace

extern "C" {

/* which shader text this library was generated from (sizes in bytes; checked by tests against nothing but itself:
 * a stale library after a reference update shows up in the build log) */
const char *svo_ref_about(void) {
  return "svotrace.comp + svobeam.comp compiled by g++ through oracle/glsl_shim.h (oracle/build_ref.py)";
}

/* One intersectOctree call of svotrace.comp (:211-432) with fresh per-invocation state.  res is in/out. */
int svo_ref_cast(const uint8_t *nodes, uint64_t nbytes, const float o[3], const float d[3], int maxDepth,
                 int coneTrace, int coneDepth, svo_o_cast_result *res, uint32_t *loop_iter) {
  std::lock_guard<std::mutex> g(g_lock);
  bind_trace(nodes, nbytes, maxDepth, coneDepth, 2);
  ref_svotrace::Invocation inv;
  ref_svotrace::Invocation::castResult r = {};
  to_shader(*res, r);
  r.depth = res->depth;
  glsl::vec3 dir = v3(d);
  bool hit = inv.intersectOctree(v3(o), dir, 1.0f / dir, r, maxDepth, coneTrace != 0);
  from_shader(r, *res);
  res->depth = r.depth;
  if (loop_iter) *loop_iter = inv.cast_log[0].loop_iter;
  return hit ? 1 : 0;
}

void svo_ref_cast_rays(const uint8_t *nodes, uint64_t nbytes, const svo_o_ray *rays, uint64_t n, int maxDepth,
                       svo_o_hit *out, int nthreads) {
  std::lock_guard<std::mutex> g(g_lock);
  bind_trace(nodes, nbytes, maxDepth, 11, 2);
  nthreads = clamp_threads(nthreads);
  rays_job *jobs = (rays_job *)calloc((size_t)nthreads, sizeof *jobs);
  for (int t = 0; t < nthreads; t++) jobs[t] = rays_job{rays, out, n, maxDepth, t, nthreads};
  run_threads(jobs, nthreads, rays_worker);
  free(jobs);
}

/* svotrace.comp main() dispatched over rows [y0, y1) of a width x height image; planes as in svo_oracle_render.
 * Returns 0, or -1 when the frame asks for something the shipped shader does not have (mirrorValue, flags). */
int svo_ref_render(const uint8_t *nodes, uint64_t nbytes, const svo_o_frame *f, int width, int height, int y0,
                   int y1, const float *beam, uint8_t *rgba8, float *depth, float *radiance, uint32_t *hit_id,
                   uint32_t *iter, float *primary_t, int nthreads) {
  if (f->mirrorValue != 0 || (f->flags & 1)) return -1;
  std::lock_guard<std::mutex> g(g_lock);
  bind_trace(nodes, nbytes, f->maxDepth, f->coneDepth, f->casts);
  using namespace ref_svotrace;
  camPos = v3(f->camPos);                                            /* Main.java:269 */
  cam[1] = v3(f->l1); cam[2] = v3(f->l2); cam[3] = v3(f->r1); cam[4] = v3(f->r2); /* :270-273 */
  frameNumber = f->frameNumber;                                      /* :276 */
  renderMode = f->renderMode;                                        /* :279 */
  useBeamOptimization = f->useBeam != 0;                             /* :283 */
  framebufferImage = glsl::image2D{width, height, glsl::IMG_RGBA8, rgba8, radiance};   /* Main.java:66-70 */
  depthbufferImage = glsl::image2D{width, height, glsl::IMG_R32F, nullptr, depth};      /* :73-77 */
  beambufferImage = glsl::image2D{width / 4, height / 4, glsl::IMG_R32F, nullptr, const_cast<float *>(beam)}; /* :79-86 */
  nthreads = clamp_threads(nthreads);
  job *jobs = (job *)calloc((size_t)nthreads, sizeof *jobs);
  for (int t = 0; t < nthreads; t++) jobs[t] = job{width, y0, y1, t, nthreads, depth, hit_id, iter, primary_t};
  run_threads(jobs, nthreads, render_worker);
  free(jobs);
  return 0;
}

/* svobeam.comp main() (:617-636) dispatched over the (width/4) x (height/4) beam image (Main.java:257-266). */
int svo_ref_beam(const uint8_t *nodes, uint64_t nbytes, const svo_o_frame *f, int width, int height,
                 float *beam_out, int nthreads) {
  (void)nthreads;
  std::lock_guard<std::mutex> g(g_lock);
  using namespace ref_svobeam;
  octreeBuffer.bytes = nodes;
  octreeBuffer.nbytes = nbytes;
  bufferEnd = (int)nbytes;
  ref_max_depth = f->maxDepth;
  camPos = v3(f->camPos);
  cam[1] = v3(f->l1); cam[2] = v3(f->l2); cam[3] = v3(f->r1); cam[4] = v3(f->r2);
  frameNumber = f->frameNumber;
  renderMode = f->renderMode;
  int bw = width / 4, bh = height / 4;
  framebufferImage = glsl::image2D{width, height, glsl::IMG_RGBA8, nullptr, nullptr};
  beambufferImage = glsl::image2D{bw, bh, glsl::IMG_R32F, nullptr, beam_out};
  for (int gy = 0; gy < bh; gy++)
    for (int gx = 0; gx < bw; gx++) {
      Invocation inv;
      inv.gl_GlobalInvocationID.xy = glsl::uvec2((uint32_t)gx, (uint32_t)gy);
      inv.main();
    }
  return 0;
}

}  // extern "C"
