/*
 * svo_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of
 * /root/reference/src/shaders/svotrace.comp (hot path) and
 * /root/reference/src/shaders/svobeam.comp:617-636 (beam pre-pass main).
 *
 * Pinned bit for bit to the reference's own shaders compiled for the CPU (oracle/_ref, see svo_oracle.h and
 * tests/test_oracle_ref.py).  Arithmetic contract: oracle_math.h.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile).
 *
 * Interpretation choices where GLSL leaves behaviour undefined (each is
 * mirrored by the CUDA validation build and listed in DESIGN.md):
 *   U1  out-of-range SSBO reads and out-of-image beam-buffer loads return 0 (robust buffer / image access);
 *   U2  `castResult res` and `matcolor` start zeroed (uninitialised upstream);
 *   U3  `octstack` starts zeroed per invocation and persists across the casts
 *       of one pixel (it is a global in the shader);
 *   U4  `out` struct parameters keep the caller's stale fields;
 *   U5  imageStore to rgba8: NaN -> 0, clamp to [0,1], floor(c*255+0.5);
 *   U6  findMSB(0) = -1 cannot occur (differing_bits != 0 whenever POP runs);
 *   U8  frame.flags bit 0 enables the progressive running mean the shader has commented out (:712-719);
 *       the rgba8 plane is then read before it is written (unorm8 -> float = c / 255).
 *   U7  hit_id / iter / primary_t planes are new outputs (the shader's
 *       pointer store is commented out, svotrace.comp:728): hit_id is
 *       res.pointer of the PRIMARY cast on hit, NO_HIT on miss; iter is the
 *       primary cast's loop-iteration count whether it hit or not.
 */
#include "svo_oracle.h"
#include "oracle_math.h"

#include <pthread.h>
#include <stdlib.h>

#define MAX_SCALE SVO_O_MAX_SCALE
#define EPSILON 3.552713678800501e-15f /* svotrace.comp:31 (2^-48) */
#define PI_F 3.14159265359f            /* svotrace.comp:33 */
#define SQRT3_F 1.73205080757f         /* svotrace.comp:34 */

typedef struct {
  const uint8_t *nodes;
  uint64_t nbytes;
} buf_t;

/* svotrace.comp:75-79 getByte; U1 */
static inline int get_byte(const buf_t *b, uint32_t p) {
  return ((uint64_t)p < b->nbytes) ? (int)b->nodes[p] : 0;
}

/* svotrace.comp:81-86 */
typedef struct {
  int value;
  int cp;
  int leafMask;
  uint32_t descriptor;
} node_t;

/* svotrace.comp:88-101 */
static node_t extract_node(const buf_t *b, uint32_t p) {
  node_t r;
  r.descriptor = p;
  r.value = get_byte(b, p);
  uint32_t c = ((uint32_t)get_byte(b, p + 1u) << 24) | ((uint32_t)get_byte(b, p + 2u) << 16) |
               ((uint32_t)get_byte(b, p + 3u) << 8) | (uint32_t)get_byte(b, p + 4u);
  r.cp = (int)c;
  r.leafMask = (get_byte(b, p + 5u) << 8) | get_byte(b, p + 6u);
  return r;
}
/* svotrace.comp:103-108 */
static node_t extract_leaf(const buf_t *b, uint32_t p) {
  node_t r = {get_byte(b, p), 0, 0, p};
  r.leafMask = get_byte(b, p + 1u) | (get_byte(b, p + 2u) << 8);
  return r;
}
/* svotrace.comp:110-114 */
static node_t extract_non_surface_leaf(const buf_t *b, uint32_t p) {
  node_t r = {get_byte(b, p), 0, 0, p};
  return r;
}
/* svotrace.comp:116-130 */
static node_t extract_subdividable_leaf(const buf_t *b, uint32_t p) {
  node_t r = extract_node(b, p);
  r.cp = 0;
  return r;
}

/* svotrace.comp:132-157 */
static node_t extract_child(const buf_t *b, uint32_t parentPointer, uint32_t childPointer,
                            uint32_t child, int leafMask, uint32_t *endPointer, int *rec_size) {
  uint32_t i = 0;
  uint32_t pointer = childPointer + parentPointer;
  while (i < child) {
    int localMask = (leafMask & (0x0003 << (i << 1))) >> (i << 1);
    if (localMask == 0 || localMask == 2) pointer += 7u;
    else if (localMask == 1) pointer += 3u;
    else if (localMask == 3) pointer += 1u;
    i++;
  }
  *endPointer = pointer;
  int localMask = (leafMask & (0x0003 << (child << 1))) >> (child << 1);
  if (localMask == 0) { *rec_size = 7; return extract_node(b, pointer); }
  if (localMask == 1) { *rec_size = 3; return extract_leaf(b, pointer); }
  if (localMask == 2) { *rec_size = 7; return extract_subdividable_leaf(b, pointer); }
  *rec_size = 1;
  return extract_non_surface_leaf(b, pointer);
}

/* svotrace.comp:199-202 */
typedef struct {
  node_t node;
  float tmax;
} stack_entry_t;

typedef struct {
  stack_entry_t s[MAX_SCALE + 1];
} stack_t;

static inline int find_msb(uint32_t v) { return v ? 31 - __builtin_clz(v) : -1; }

/* svotrace.comp:211-432.  Returns 1 on hit; on miss -(loop iterations) - 1, so that callers can fill the iter
 * plane (U7) without a castResult field the shader does not have.  `st` is the invocation's octstack (U3). */
static int intersect_octree(const buf_t *b, stack_t *st, const float origin[3], const float dir_in[3],
                            svo_o_cast_result *res, int maxDepth, int coneTrace, int coneDepth,
                            svo_o_stats *stats) {
  float dir[3] = {dir_in[0], dir_in[1], dir_in[2]};
  uint32_t written = 0; /* instrumentation only */
  if (stats) { stats->casts++; stats->record_bytes += 7; }

  res->debugColor[0] = 0.3f; res->debugColor[1] = 0.3f; res->debugColor[2] = 0.6f; /* :213 */
  node_t parent = extract_node(b, 0u);                                              /* :222 */
  uint32_t iter = 0;

  for (int a = 0; a < 3; a++)                                                       /* :226-228 */
    if (om_abs(dir[a]) < EPSILON) dir[a] = EPSILON * om_sign(dir[a]);

  float t_coef[3], t_bias[3];
  for (int a = 0; a < 3; a++) {                                                     /* :230-236 */
    t_coef[a] = 1.0f / -om_abs(dir[a]);
    t_bias[a] = t_coef[a] * origin[a];
  }
  uint32_t octant_mask = 0;                                                         /* :238-241 */
  for (int a = 0; a < 3; a++)
    if (dir[a] > 0.0f) {
      octant_mask ^= (1u << a);
      float t3 = 3.0f * t_coef[a];
      t_bias[a] = t3 - t_bias[a];
    }

  float lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    float c2 = 2.0f * t_coef[a];
    lo[a] = c2 - t_bias[a];
    hi[a] = t_coef[a] - t_bias[a];
  }
  float t_min = om_max(om_max(lo[0], lo[1]), lo[2]);                                /* :243 */
  float t_max = om_min(om_min(hi[0], hi[1]), hi[2]);                                /* :244 */
  t_min = om_max(t_min, 0.0f);                                                      /* :245 */
  float h = t_max;                                                                  /* :247 */

  uint32_t idx = 0;
  float pos[3] = {1.0f, 1.0f, 1.0f};
  int scale = MAX_SCALE - 1;
  float scale_exp2 = 0.5f;
  int child_descriptor = 0;

  for (int a = 0; a < 3; a++) {                                                     /* :255-257 */
    float c15 = 1.5f * t_coef[a];
    if (c15 - t_bias[a] > t_min) { idx ^= (1u << a); pos[a] = 1.5f; }
  }
  uint32_t child_shift = 0;

  while (scale < MAX_SCALE) {                                                       /* :262 */
    iter++;
    if (iter > SVO_O_MAX_RAYCAST_ITERATIONS) {                                      /* :264-266 */
      if (stats) { stats->capped++; stats->iters += iter - 1; }
      return -(int)iter - 1; /* miss; carries the loop count (1501) for U7 */
    }
    if (child_descriptor == 0) child_descriptor = parent.cp;                        /* :267-269 */
    if (t_min > 0.05f && coneTrace) maxDepth = coneDepth;                           /* :275-277 */

    float corner[3];
    for (int a = 0; a < 3; a++) {                                                   /* :280-282 */
      float m = pos[a] * t_coef[a];
      corner[a] = m - t_bias[a];
    }
    float tc_max = om_min(om_min(corner[0], corner[1]), corner[2]);                 /* :283 */

    child_shift = idx ^ octant_mask;                                                /* :286 */
    int rec_size;
    node_t child = extract_child(b, parent.descriptor, (uint32_t)child_descriptor, child_shift,
                                 parent.leafMask, &res->pointer, &rec_size);        /* :294 */
    if (stats) stats->record_bytes += (uint64_t)rec_size;

    if (child.value != 0 && t_min <= t_max) {                                       /* :295 */
      if (MAX_SCALE - scale == maxDepth) break;                                     /* :300-302 */
      float tv_max = om_min(t_max, tc_max);                                         /* :304 */
      float one_half = scale_exp2 * 0.5f;
      float center[3];
      for (int a = 0; a < 3; a++) {                                                 /* :306-308 */
        float m = one_half * t_coef[a];
        center[a] = m + corner[a];
      }
      if (t_min <= tv_max) {                                                        /* :310 */
        if (child.cp == 0) break;                                                   /* :311-313 */
        if (tc_max < h) {                                                           /* :316-319 PUSH */
          st->s[scale].node = parent;
          st->s[scale].tmax = t_max;
          written |= 1u << scale;
        }
        h = tc_max;
        parent = child;
        idx = 0u;
        --scale;
        scale_exp2 = one_half;
        for (int a = 0; a < 3; a++)                                                 /* :328-330 */
          if (center[a] > t_min) { idx ^= (1u << a); pos[a] += scale_exp2; }
        t_max = tv_max;
        child_descriptor = 0;
        continue;
      }
    }
    /* ADVANCE :337-344 */
    uint32_t step_mask = 0u;
    for (int a = 0; a < 3; a++)
      if (corner[a] <= tc_max) { step_mask ^= (1u << a); pos[a] -= scale_exp2; }
    t_min = tc_max;
    idx ^= step_mask;

    if ((idx & step_mask) != 0) {                                                   /* POP :347-368 */
      uint32_t differing_bits = 0;
      for (int a = 0; a < 3; a++)
        if ((step_mask & (1u << a)) != 0)
          differing_bits |= om_f2u(pos[a]) ^ om_f2u(pos[a] + scale_exp2);
      scale = find_msb(differing_bits);
      scale_exp2 = om_u2f(((uint32_t)scale - (uint32_t)MAX_SCALE + 127u) << 23);
      int sidx = scale;
      if (sidx < 0) sidx = 0;                 /* U6: unreachable */
      if (sidx > MAX_SCALE) sidx = MAX_SCALE; /* differing bit above 23 (pos left [1,2)): miss exit anyway */
      if (stats && scale < MAX_SCALE && !(written & (1u << sidx))) stats->stale_pops++;
      parent = st->s[sidx].node;
      t_max = st->s[sidx].tmax;
      uint32_t sh[3];
      for (int a = 0; a < 3; a++) {
        sh[a] = (scale >= 0 && scale < 32) ? (om_f2u(pos[a]) >> scale) : 0u;
        pos[a] = om_u2f((scale >= 0 && scale < 32) ? (sh[a] << scale) : 0u);
      }
      idx = (sh[0] & 1u) | ((sh[1] & 1u) << 1) | ((sh[2] & 1u) << 2);
      h = 0.0f;
      child_descriptor = 0;
    }
  }
  if (stats) stats->iters += iter;

  if (scale >= MAX_SCALE) {                                                         /* :371-377 */
    float it = 0.01f * (float)iter;
    res->debugColor[0] = it; res->debugColor[1] = it; res->debugColor[2] = it;
    return -(int)iter - 1; /* negative = miss; carries the loop count for U7 (res->iter untouched) */
  }

  float norm[3] = {0.0f, 0.0f, 0.0f};                                               /* :380 */
  int rec_size;
  node_t target = extract_child(b, parent.descriptor, (uint32_t)child_descriptor, child_shift,
                                parent.leafMask, &res->pointer, &rec_size);         /* :381 */
  if (target.leafMask != 0) {                                                       /* :382-388 */
    int raw = target.leafMask;
    float n3[3];
    n3[0] = (float)((raw % 10) - 5);
    n3[1] = (float)((((raw % 100) - (raw % 10)) / 10) - 5);
    n3[2] = (float)(((raw - (raw % 100)) / 100) - 5);
    om_normalize3(n3, norm);
  }
  res->t = t_min;                                                                   /* :403-408 */
  res->value = (uint32_t)target.value;
  res->iter = iter;
  res->scale = scale_exp2;
  res->depth = (uint32_t)(MAX_SCALE - scale);
  for (int a = 0; a < 3; a++) {
    res->normal[a] = norm[a];
    float td = t_min * dir[a];                                                      /* :410 */
    float ns = norm[a] * scale_exp2;
    ns = ns * 2.0f;
    float od = origin[a] + td;
    res->hitPos[a] = od + ns;
    float vp = pos[a];                                                              /* :413-421 */
    if (dir[a] > 0.0f) { vp = 3.0f - vp; vp = vp - scale_exp2; }
    float nv = norm[a] * scale_exp2;
    nv = nv * 2.0f;
    nv = nv * 1.74f;
    res->voxelPos[a] = vp + nv;
  }
  float dc = 0.005f * (float)iter;                                                  /* :428 */
  res->debugColor[0] = dc; res->debugColor[1] = dc; res->debugColor[2] = dc;
  if (stats) stats->hits++;
  return (scale < MAX_SCALE && t_min <= t_max) ? 1 : -(int)iter - 1;                /* :431 */
}

/* public single cast */
int svo_oracle_cast(const uint8_t *nodes, uint64_t nbytes, const float o[3], const float d[3],
                    int maxDepth, int coneTrace, int coneDepth, svo_o_cast_result *res,
                    svo_o_stats *stats) {
  buf_t b = {nodes, nbytes};
  stack_t st;
  memset(&st, 0, sizeof st);
  int r = intersect_octree(&b, &st, o, d, res, maxDepth, coneTrace, coneDepth, stats);
  return r > 0 ? 1 : 0;
}

/* svotrace.comp:26-29 */
static float rand2(float x, float y) {
  float a = x * 12.9898f;
  float c = y * 78.233f;
  float s = om_sin(a + c);
  return om_fract(s * 43758.5453f);
}
float svo_oracle_rand(float x, float y) { return rand2(x, y); }
float svo_oracle_sin(float x) { return om_sin(x); }
float svo_oracle_cos(float x) { return om_cos(x); }
float svo_oracle_acos(float x) { return om_acos(x); }
float svo_oracle_exp(float x) { return om_exp(x); }

static void matcolor_table(uint32_t value, float mc[3]) { /* :514-522 / :578-586 */
  if (value == 1) { mc[0] = 0.84f; mc[1] = 0.86f; mc[2] = 0.78f; }
  if (value == 2) { mc[0] = 0.57f; mc[1] = 0.5f; mc[2] = 0.31f; }
  if (value == 3) { mc[0] = 0.37f; mc[1] = 0.43f; mc[2] = 0.27f; }
}

typedef struct {
  float color[3];
  float depth;
  uint32_t hit_id;
  uint32_t iter;
  float primary_t;
} px_out_t;

/* svotrace.comp:435-646 trace() */
static void trace(const buf_t *b, const svo_o_frame *f, float beamDist, const float origin_in[3],
                  const float dir_in[3], float seed0, float seed1, float seed2, px_out_t *out,
                  svo_o_stats *stats) {
  svo_o_cast_result res;
  memset(&res, 0, sizeof res); /* U2 */
  res.t = 2.0f;                /* :437 */
  stack_t st;
  memset(&st, 0, sizeof st);   /* U3 */
  float origin[3], dir[3];
  for (int a = 0; a < 3; a++) {
    dir[a] = dir_in[a];
    float m = dir_in[a] * beamDist;
    origin[a] = origin_in[a] + m; /* :438 */
  }
  float accum[3] = {0, 0, 0}, mask[3] = {1, 1, 1};
  const float inv_sqrt3 = 1.0f / sqrtf(3.0f); /* normalize(vec3(1)) : (1+1)+1 = 3 */
  const float sun_dir[3] = {inv_sqrt3, inv_sqrt3, inv_sqrt3};
  int mode = f->renderMode;
  int maxDepth = f->maxDepth;

  if (mode == 0) {                                                                  /* :443-560 */
    int intersect = 1;
    for (int i = 0; i < f->casts; i++) {
      int coneTrace = (i != 0);
      int r = intersect_octree(b, &st, origin, dir, &res, maxDepth, coneTrace, f->coneDepth, stats);
      intersect = r > 0;
      if (i == 0) {
        out->iter = intersect ? res.iter : (uint32_t)(-r - 1);
        out->hit_id = intersect ? res.pointer : SVO_O_NO_HIT;
        out->primary_t = intersect ? res.t : 0.0f;
      }
      if (!intersect && i == 0) {                                                   /* :448-452 */
        accum[0] += 0.6725f - dir[1] * 0.4f;
        accum[1] += 0.8784f - dir[1] * 0.4f;
        accum[2] += 1.0f - dir[1] * 0.25f;
        break;
      }
      float normal[3] = {res.normal[0], res.normal[1], res.normal[2]};              /* :476 */
      float hitpoint[3] = {res.voxelPos[0], res.voxelPos[1], res.voxelPos[2]};      /* :481 */
      float ra = rand2(seed0, seed2 * 0.1f);                                        /* :486 */
      float rb = rand2(seed1, seed2 * 0.02f);
      float rnd = rand2(seed0 + ra, seed1 + rb);
      float rand1 = 2.0f * PI_F;                                                    /* :487 */
      rand1 = rand1 * rnd;
      const float *w = normal;                                                      /* :494-497 */
      float axis[3] = {1.0f, 0.0f, 0.0f};
      if (om_abs(w[0]) > 0.1f) { axis[0] = 0.0f; axis[1] = 1.0f; }
      float cr[3], u[3], v[3], newdir[3];
      om_cross3(axis, w, cr);
      om_normalize3(cr, u);
      om_cross3(w, u, v);
      if (f->mirrorValue != 0 && res.value == (uint32_t)f->mirrorValue) {           /* :500-504 (extension) */
        float dn = om_dot3(dir, normal);
        dn = 2.0f * dn;
        for (int a = 0; a < 3; a++) { float m = dn * normal[a]; newdir[a] = dir[a] - m; }
      } else {                                                                      /* :506 */
        float c = om_cos(rand1), s = om_sin(rand1);
        float omr = 1.0f - rnd;
        float nd[3];
        for (int a = 0; a < 3; a++) {
          float uc = u[a] * c;
          float vs = v[a] * s;
          float ww = w[a] * omr;
          float sum = uc + vs;
          nd[a] = sum + ww;
        }
        om_normalize3(nd, newdir);
      }
      for (int a = 0; a < 3; a++) { origin[a] = hitpoint[a]; dir[a] = newdir[a]; } /* :508-509 */
      float matcolor[3] = {hitpoint[0] - 1.0f, hitpoint[1] - 1.0f, hitpoint[2] - 1.0f}; /* :511 */
      matcolor_table(res.value, matcolor);
      if (intersect) {                                                              /* :531-535 */
        out->depth = res.t;
        float dnn = om_dot3(newdir, normal);
        for (int a = 0; a < 3; a++) {
          float e = mask[a] * 0.0f; /* mask * matemi (NaN-propagating) */
          accum[a] = accum[a] + e;
          mask[a] = mask[a] * matcolor[a];
          mask[a] = mask[a] * dnn;
        }
      } else {                                                                      /* :536-557 */
        float diff = om_acos(om_dot3(dir, sun_dir));
        if (diff < 0.4f)
          for (int a = 0; a < 3; a++) { float e = mask[a] * 7.0f; accum[a] = accum[a] + e; }
        for (int a = 0; a < 3; a++) { float e = mask[a] * 1.0f; accum[a] = accum[a] + e; }
        out->depth = 0.0f;
        break;
      }
    }
    for (int a = 0; a < 3; a++) out->color[a] = accum[a];
    return;
  }

  /* modes 1,2,3 share one primary cast */
  if (mode == 1 || mode == 2 || mode == 3) {
    int r = intersect_octree(b, &st, origin, dir, &res, maxDepth, 0, f->coneDepth, stats);
    int hit = r > 0;
    out->iter = hit ? res.iter : (uint32_t)(-r - 1);
    out->hit_id = hit ? res.pointer : SVO_O_NO_HIT;
    out->primary_t = hit ? res.t : 0.0f;
    if (mode == 1) {                                                                /* :561-571 */
      out->depth = hit ? res.t : 0.0f;
      for (int a = 0; a < 3; a++) out->color[a] = res.debugColor[a];
      return;
    }
    if (mode == 3) {                                                                /* :633-642 */
      if (hit) {
        out->depth = res.t;
        for (int a = 0; a < 3; a++) { float m = res.normal[a] * 0.5f; out->color[a] = m + 0.5f; }
      } else {
        out->depth = 0.0f;
        out->color[0] = out->color[1] = out->color[2] = 0.0f;
      }
      return;
    }
    /* mode 2 :572-632 */
    if (hit) {
      out->depth = res.t;
      float matcolor[3] = {0, 0, 0}; /* U2 */
      matcolor_table(res.value, matcolor);
      const float sd = 0.5f / sqrtf(0.75f); /* normalize(vec3(0.5)): (0.25+0.25)+0.25 */
      const float sun2[3] = {sd, sd, sd};
      if (res.depth >= 10) {                                                        /* :588-593 */
        float ph = om_dot3(res.normal, sun2);
        ph = ph * 0.1f;
        for (int a = 0; a < 3; a++) matcolor[a] = matcolor[a] + ph;
      } else {
        const float up[3] = {0.0f, 1.0f, 0.0f};
        float ph = om_dot3(up, sun2);
        ph = ph * 0.1f;
        for (int a = 0; a < 3; a++) matcolor[a] = matcolor[a] + ph;
      }
      float trueDist = res.t + beamDist;                                            /* :595-598 */
      float base = -0.5f * trueDist;
      float lambdag = om_exp(base * 2.0f);
      float lambdab = om_exp(base * 4.0f);
      float lambdar = om_exp(base * 1.0f);
      {                                                                             /* :602-604, fog = 1 */
        float a0 = lambdar * matcolor[0], b0 = (1.0f - lambdar) * 1.0f;
        matcolor[0] = a0 + b0;
        float a1 = lambdag * matcolor[1], b1 = (1.0f - lambdag) * 1.0f;
        matcolor[1] = a1 + b1;
        float a2 = lambdab * matcolor[2], b2 = (1.0f - lambdab) * 1.0f;
        matcolor[2] = a2 + b2;
      }
      float sorigin[3] = {res.voxelPos[0], res.voxelPos[1], res.voxelPos[2]};
      int sr = intersect_octree(b, &st, sorigin, sun2, &res, maxDepth, 0, f->coneDepth, stats); /* :607 */
      if (sr > 0 && res.t > res.scale * SQRT3_F) {
        for (int a = 0; a < 3; a++) matcolor[a] = matcolor[a] - 0.2f;
      } else if (res.iter > 260) {                                                  /* :616-619 */
        float pen = 0.05f * (float)res.iter;
        pen = pen / 100.0f;
        for (int a = 0; a < 3; a++) matcolor[a] = matcolor[a] - pen;
      }
      for (int a = 0; a < 3; a++) out->color[a] = matcolor[a];
    } else {                                                                        /* :626-632 */
      out->depth = 0.0f;
      out->color[0] = 0.6725f - dir[1] * 0.4f;
      out->color[1] = 0.8784f - dir[1] * 0.4f;
      out->color[2] = 1.0f - dir[1] * 0.25f;
    }
    return;
  }
  /* mode 4 (:643-645) returns the uninitialised res.voxelPos: zero under U2 */
  for (int a = 0; a < 3; a++) out->color[a] = res.voxelPos[a];
}

static inline uint8_t quant8(float c) { /* U5 */
  if (c != c) return 0;
  if (c < 0.0f) c = 0.0f;
  if (c > 1.0f) c = 1.0f;
  return (uint8_t)floorf(c * 255.0f + 0.5f);
}

typedef struct {
  buf_t b;
  const svo_o_frame *f;
  int width, height, y0, y1, tid, nthreads;
  const float *beam;
  uint8_t *rgba8;
  float *depth, *radiance, *primary_t;
  uint32_t *hit_id, *iter;
  svo_o_stats stats;
} render_job_t;

/* svotrace.comp:649-729 main() for one pixel */
static void shade_pixel(render_job_t *j, int x, int y) {
  const svo_o_frame *f = j->f;
  float beamDist = 0.0f;
  /* :656-658.  The beam image is (W/4) x (H/4) texels (Main.java:82-83, integer division); imageLoad outside an image
   * returns 0 (GL robust image access), which happens for the last rows/columns when W or H is not a multiple of 4 */
  if (f->useBeam && j->beam && x / 4 < j->width / 4 && y / 4 < j->height / 4)
    beamDist = j->beam[(size_t)(y / 4) * (size_t)(j->width / 4) + (size_t)(x / 4)];
  float px = ((float)x + 0.5f) / (float)j->width;                                   /* :662 */
  float py = ((float)y + 0.5f) / (float)j->height;
  float dir[3], nd[3];
  for (int a = 0; a < 3; a++) {                                                     /* :664 */
    float l = om_mix(f->l1[a], f->l2[a], py);
    float r = om_mix(f->r1[a], f->r2[a], py);
    dir[a] = om_mix(l, r, px);
  }
  om_normalize3(dir, nd);                                                           /* :675 */
  px_out_t o;
  memset(&o, 0, sizeof o);
  o.depth = -1.0f;                                                                  /* :672 */
  o.hit_id = SVO_O_NO_HIT;
  trace(&j->b, f, beamDist, f->camPos, nd, (float)x, (float)y, (float)f->frameNumber, &o, &j->stats);
  if (x < 10 && y < 10) {                                                           /* :696-700 */
    int first_zero = (get_byte(&j->b, 0) | get_byte(&j->b, 1) | get_byte(&j->b, 2) | get_byte(&j->b, 3)) == 0;
    o.color[0] = 1.0f;
    o.color[1] = first_zero ? 0.0f : 1.0f;
    o.color[2] = first_zero ? 0.0f : 1.0f;
  }
  size_t p = (size_t)y * (size_t)j->width + (size_t)x;
  if ((f->flags & 1) && j->rgba8 && f->frameNumber > 1) {                           /* :712-719 (commented out upstream) */
    for (int a = 0; a < 3; a++) {
      float last = (float)j->rgba8[4 * p + a] / 255.0f;                             /* imageLoad of the rgba8 image */
      if (f->frameNumber < 100) {                                                   /* MAX_FRAME_ITER :43 */
        float s = (float)f->frameNumber * last;
        s = s + o.color[a];
        o.color[a] = s / (float)(f->frameNumber + 1);
      } else {
        o.color[a] = last;
      }
    }
  }
  if (j->rgba8) {                                                                   /* :726 */
    j->rgba8[4 * p + 0] = quant8(o.color[0]);
    j->rgba8[4 * p + 1] = quant8(o.color[1]);
    j->rgba8[4 * p + 2] = quant8(o.color[2]);
    j->rgba8[4 * p + 3] = 255;
  }
  if (j->depth) j->depth[p] = o.depth;                                              /* :727 */
  if (j->radiance) {
    j->radiance[4 * p + 0] = o.color[0];
    j->radiance[4 * p + 1] = o.color[1];
    j->radiance[4 * p + 2] = o.color[2];
    j->radiance[4 * p + 3] = 1.0f;
  }
  if (j->hit_id) j->hit_id[p] = o.hit_id;
  if (j->iter) j->iter[p] = o.iter;
  if (j->primary_t) j->primary_t[p] = o.primary_t;
}

static void *render_worker(void *arg) {
  render_job_t *j = (render_job_t *)arg;
  for (int y = j->y0 + j->tid; y < j->y1; y += j->nthreads)
    for (int x = 0; x < j->width; x++) shade_pixel(j, x, y);
  return NULL;
}

static void stats_add(svo_o_stats *d, const svo_o_stats *s) {
  d->casts += s->casts; d->iters += s->iters; d->record_bytes += s->record_bytes;
  d->stale_pops += s->stale_pops; d->capped += s->capped; d->hits += s->hits;
}

void svo_oracle_render(const uint8_t *nodes, uint64_t nbytes, const svo_o_frame *frame, int width,
                       int height, int y0, int y1, const float *beam, uint8_t *rgba8, float *depth,
                       float *radiance, uint32_t *hit_id, uint32_t *iter, float *primary_t,
                       int nthreads, svo_o_stats *stats) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  render_job_t *jobs = (render_job_t *)calloc((size_t)nthreads, sizeof *jobs);
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
  for (int t = 0; t < nthreads; t++) {
    render_job_t *j = &jobs[t];
    j->b.nodes = nodes; j->b.nbytes = nbytes; j->f = frame;
    j->width = width; j->height = height; j->y0 = y0; j->y1 = y1; j->tid = t; j->nthreads = nthreads;
    j->beam = beam; j->rgba8 = rgba8; j->depth = depth; j->radiance = radiance;
    j->hit_id = hit_id; j->iter = iter; j->primary_t = primary_t;
    if (nthreads > 1) pthread_create(&th[t], NULL, render_worker, j);
  }
  if (nthreads == 1) render_worker(&jobs[0]);
  else for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  if (stats) for (int t = 0; t < nthreads; t++) stats_add(stats, &jobs[t].stats);
  free(jobs);
  free(th);
}

/* ---- ray stream ---------------------------------------------------------- */
typedef struct {
  buf_t b;
  const svo_o_ray *rays;
  svo_o_hit *out;
  uint64_t n;
  int maxDepth, tid, nthreads;
  svo_o_stats stats;
} rays_job_t;

static void *rays_worker(void *arg) {
  rays_job_t *j = (rays_job_t *)arg;
  uint64_t per = (j->n + (uint64_t)j->nthreads - 1) / (uint64_t)j->nthreads;
  uint64_t lo = per * (uint64_t)j->tid, hi = lo + per;
  if (hi > j->n) hi = j->n;
  for (uint64_t i = lo; i < hi; i++) {
    svo_o_cast_result res;
    memset(&res, 0, sizeof res);
    stack_t st;
    memset(&st, 0, sizeof st);
    int r = intersect_octree(&j->b, &st, j->rays[i].o, j->rays[i].d, &res, j->maxDepth, 0, 11, &j->stats);
    if (r > 0) {
      j->out[i].id = res.pointer; j->out[i].t = res.t; j->out[i].value = res.value; j->out[i].iter = res.iter;
    } else {
      j->out[i].id = SVO_O_NO_HIT; j->out[i].t = 0.0f; j->out[i].value = 0; j->out[i].iter = (uint32_t)(-r - 1);
    }
  }
  return NULL;
}

void svo_oracle_cast_rays(const uint8_t *nodes, uint64_t nbytes, const svo_o_ray *rays, uint64_t n,
                          int maxDepth, svo_o_hit *out, int nthreads, svo_o_stats *stats) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  rays_job_t *jobs = (rays_job_t *)calloc((size_t)nthreads, sizeof *jobs);
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
  for (int t = 0; t < nthreads; t++) {
    rays_job_t *j = &jobs[t];
    j->b.nodes = nodes; j->b.nbytes = nbytes; j->rays = rays; j->out = out; j->n = n;
    j->maxDepth = maxDepth; j->tid = t; j->nthreads = nthreads;
    if (nthreads > 1) pthread_create(&th[t], NULL, rays_worker, j);
  }
  if (nthreads == 1) rays_worker(&jobs[0]);
  else for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  if (stats) for (int t = 0; t < nthreads; t++) stats_add(stats, &jobs[t].stats);
  free(jobs);
  free(th);
}

/* ---- beam pre-pass (svobeam.comp:617-636) --------------------------------
 * One UN-normalised ray through pixel (4gx, 4gy); stores res.t.  Only the
 * traversal differences that affect res.t are kept: coneTrace = true with
 * the LOD cut commented out upstream (svobeam.comp:261-267) -> no cut.
 * res.t is undefined on miss upstream; 0 here. */
void svo_oracle_beam(const uint8_t *nodes, uint64_t nbytes, const svo_o_frame *f, int width,
                     int height, float *beam_out, int nthreads) {
  (void)nthreads;
  buf_t b = {nodes, nbytes};
  int bw = width / 4, bh = height / 4;
  for (int gy = 0; gy < bh; gy++)
    for (int gx = 0; gx < bw; gx++) {
      float px = ((float)(gx * 4) + 0.5f) / (float)width;
      float py = ((float)(gy * 4) + 0.5f) / (float)height;
      float dir[3];
      for (int a = 0; a < 3; a++) {
        float l = om_mix(f->l1[a], f->l2[a], py);
        float r = om_mix(f->r1[a], f->r2[a], py);
        dir[a] = om_mix(l, r, px);
      }
      svo_o_cast_result res;
      memset(&res, 0, sizeof res);
      stack_t st;
      memset(&st, 0, sizeof st);
      int r = intersect_octree(&b, &st, f->camPos, dir, &res, f->maxDepth, 0, f->coneDepth, NULL);
      beam_out[(size_t)gy * (size_t)bw + (size_t)gx] = r > 0 ? res.t : 0.0f;
    }
}
