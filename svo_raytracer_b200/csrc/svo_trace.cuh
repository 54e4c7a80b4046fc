// svo_trace.cuh -- device code of the SVO trace path for sm_100a.
//
// What it computes is defined by the reference compute shader
// src/shaders/svotrace.comp (intersectOctree :211-432, trace :435-646, main
// :649-729); how it computes it is new:
//
//  * The reference walks 7/3/1-byte unaligned records and fetches one child
//    record per loop iteration (extractChild :132-157: a <=7-trip scan plus
//    1..7 getByte word loads).  Here svo_upload transcodes the stream once
//    into one 8-byte descriptor per interior node,
//        .x = index of the node's first interior child (BFS order, so the
//             upper levels are a prefix of the array),
//        .y = leafMask type codes [0:16) | child value!=0 mask [16:24) |
//             child-has-descriptor mask [24:32),
//    so an iteration over an empty or leaf child touches no memory at all and
//    a PUSH is one aligned 8-byte read-only load (ld.global.nc).  The original
//    bytes stay on the device and are read once per cast, at the hit, to
//    produce value / packed normal; the hit id is the reference's byte offset
//    (res.pointer), recomputed from refbase[parent] + a popcount prefix of the
//    type codes.
//  * The control flow (iteration count, PUSH/ADVANCE/POP order, the 1500
//    iteration cap, the sticky cone LOD cut, every quirk listed in DESIGN.md)
//    is reproduced exactly; Ops<false> rounds every operation separately
//    (validation build semantics of --fmad=false, independent of compiler
//    flags), Ops<true> lets the t-arithmetic contract into FFMA.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "detmath.cuh"
#include "svo_kernels.h"

namespace svo {

constexpr int kMaxScale = 23;           // svotrace.comp:39
constexpr int kMaxIterations = 1500;    // svotrace.comp:41
constexpr uint32_t kNoHit = 0xFFFFFFFFu;

template <bool FAST>
struct Ops;
template <>
struct Ops<false> {
  static SVO_DI float mul(float a, float b) { return __fmul_rn(a, b); }
  static SVO_DI float sub(float a, float b) { return __fsub_rn(a, b); }
  static SVO_DI float msub(float a, float b, float c) { return __fsub_rn(__fmul_rn(a, b), c); }  // a*b - c
  static SVO_DI float madd(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }  // a*b + c
};
template <>
struct Ops<true> {
  static SVO_DI float mul(float a, float b) { return a * b; }
  static SVO_DI float sub(float a, float b) { return a - b; }
  static SVO_DI float msub(float a, float b, float c) { return fmaf(a, b, -c); }
  static SVO_DI float madd(float a, float b, float c) { return fmaf(a, b, c); }
};

// castResult (svotrace.comp:186-197) minus hitPos, which nothing reads.
struct CastRes {
  uint32_t value, pointer, iter, depth;
  float t, scale;
  vec3 normal, voxelPos;
  float dbg;     // debugColor is always a grey (dbg,dbg,dbg) ...
  int dbg_init;  // ... except the (0.3,0.3,0.6) set on entry (:213), kept by the iteration-cap exit
};

SVO_DI uint32_t raw_byte(const SceneView &sc, uint32_t p) {  // getByte (:75-79); out of range reads 0
  return ((uint64_t)p < sc.nbytes) ? (uint32_t)__ldg(sc.raw + p) : 0u;
}

// byte offset of child `c` inside a sibling block whose type codes are `codes`
// (the loop of extractChild :135-145 as two popcounts)
SVO_DI uint32_t child_offset(uint32_t codes, uint32_t c) {
  uint32_t below = (1u << (2u * c)) - 1u;
  uint32_t lo = codes & 0x5555u & below;         // bit0 of each code below c
  uint32_t hi = (codes >> 1) & 0x5555u & below;  // bit1
  uint32_t n1 = __popc(lo & ~hi);                // code 1: 3 bytes
  uint32_t n3 = __popc(lo & hi);                 // code 3: 1 byte
  uint32_t n7 = c - n1 - n3;                     // codes 0,2: 7 bytes
  return 7u * n7 + 3u * n1 + n3;
}

SVO_DI float sign_glsl(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// intersectOctree (svotrace.comp:211-432).  Returns hit; `loops` = iterations run.
struct RayStats {  // per-thread counters of the instrumented build (SVO_OPT_STATS)
  uint32_t casts, iters, record_bytes;
};

template <bool FAST, bool STATS = false>
__device__ __forceinline__ bool cast_ray(const SceneView &sc, const vec3 o, vec3 d, int maxDepth,
                                         const bool coneTrace, const int coneDepth, CastRes &res,
                                         uint32_t &loops, RayStats *rs = nullptr) {
  typedef Ops<FAST> M;
  const float kEps = 3.552713678800501e-15f;  // :31
  res.dbg_init = 1;                            // :213

  if (fabsf(d.x) < kEps) d.x = fmul(kEps, sign_glsl(d.x));  // :226-228
  if (fabsf(d.y) < kEps) d.y = fmul(kEps, sign_glsl(d.y));
  if (fabsf(d.z) < kEps) d.z = fmul(kEps, sign_glsl(d.z));

  const float tx_coef = fdiv(1.0f, -fabsf(d.x));  // :230-232
  const float ty_coef = fdiv(1.0f, -fabsf(d.y));
  const float tz_coef = fdiv(1.0f, -fabsf(d.z));
  float tx_bias = M::mul(tx_coef, o.x);  // :234-236
  float ty_bias = M::mul(ty_coef, o.y);
  float tz_bias = M::mul(tz_coef, o.z);

  uint32_t octant_mask = 0;  // :238-241
  if (d.x > 0.0f) { octant_mask ^= 1u; tx_bias = M::msub(3.0f, tx_coef, tx_bias); }
  if (d.y > 0.0f) { octant_mask ^= 2u; ty_bias = M::msub(3.0f, ty_coef, ty_bias); }
  if (d.z > 0.0f) { octant_mask ^= 4u; tz_bias = M::msub(3.0f, tz_coef, tz_bias); }

  float t_min = fmaxf(fmaxf(M::msub(2.0f, tx_coef, tx_bias), M::msub(2.0f, ty_coef, ty_bias)),
                      M::msub(2.0f, tz_coef, tz_bias));                                        // :243
  float t_max = fminf(fminf(M::sub(tx_coef, tx_bias), M::sub(ty_coef, ty_bias)), M::sub(tz_coef, tz_bias));  // :244
  t_min = fmaxf(t_min, 0.0f);  // :245
  float h = t_max;             // :247

  uint32_t idx = 0;
  float px = 1.0f, py = 1.0f, pz = 1.0f;
  int scale = kMaxScale - 1;
  float scale_exp2 = 0.5f;
  if (M::msub(1.5f, tx_coef, tx_bias) > t_min) { idx ^= 1u; px = 1.5f; }  // :255-257
  if (M::msub(1.5f, ty_coef, ty_bias) > t_min) { idx ^= 2u; py = 1.5f; }
  if (M::msub(1.5f, tz_coef, tz_bias) > t_min) { idx ^= 4u; pz = 1.5f; }

  uint32_t pidx = 0;                 // parent = root (:222)
  uint2 pd = __ldg(sc.desc);         // its descriptor
  uint32_t stk_idx[kMaxScale + 1];   // octstack (:199-202): parent index + t_max per scale
  float stk_tmax[kMaxScale + 1];
  uint32_t iter = 0;
  uint32_t child_shift = 0;
  bool hit = false;

  while (scale < kMaxScale) {  // :262
    iter++;
    if (iter > (uint32_t)kMaxIterations) break;          // :264-266 (miss, debugColor stays at its entry value)
    if (t_min > 0.05f && coneTrace) maxDepth = coneDepth;  // :275-277

    const float tx_corner = M::msub(px, tx_coef, tx_bias);  // :280-283
    const float ty_corner = M::msub(py, ty_coef, ty_bias);
    const float tz_corner = M::msub(pz, tz_coef, tz_bias);
    const float tc_max = fminf(fminf(tx_corner, ty_corner), tz_corner);

    child_shift = idx ^ octant_mask;  // :286
    if (STATS) {  // size of the child record the reference fetches here (extractChild :294)
      const uint32_t code = (pd.y >> (2u * child_shift)) & 3u;
      rs->record_bytes += code == 1u ? 3u : (code == 3u ? 1u : 7u);
    }
    // child.value != 0 (:295) is bit 16+child of the parent's descriptor
    if (((pd.y >> (16u + child_shift)) & 1u) != 0u && t_min <= t_max) {
      if (kMaxScale - scale == maxDepth) { hit = true; break; }  // :300-302
      const float tv_max = fminf(t_max, tc_max);                 // :304
      const float half = M::mul(scale_exp2, 0.5f);
      const float tx_center = M::madd(half, tx_coef, tx_corner);  // :306-308
      const float ty_center = M::madd(half, ty_coef, ty_corner);
      const float tz_center = M::madd(half, tz_coef, tz_corner);
      if (t_min <= tv_max) {  // :310
        const uint32_t dmask = pd.y >> 24;
        if (((dmask >> child_shift) & 1u) == 0u) { hit = true; break; }  // child.cp == 0 (:311-313)
        if (tc_max < h) {  // PUSH :316-319
          stk_idx[scale] = pidx;
          stk_tmax[scale] = t_max;
        }
        h = tc_max;
        pidx = pd.x + __popc(dmask & ((1u << child_shift) - 1u));  // parent = child (:322)
        pd = __ldg(sc.desc + pidx);
        idx = 0u;
        --scale;
        scale_exp2 = half;
        if (tx_center > t_min) { idx ^= 1u; px = fadd(px, scale_exp2); }  // :328-330 (exact adds)
        if (ty_center > t_min) { idx ^= 2u; py = fadd(py, scale_exp2); }
        if (tz_center > t_min) { idx ^= 4u; pz = fadd(pz, scale_exp2); }
        t_max = tv_max;
        continue;
      }
    }
    // ADVANCE :337-344
    uint32_t step_mask = 0u;
    if (tx_corner <= tc_max) { step_mask ^= 1u; px = fsub(px, scale_exp2); }
    if (ty_corner <= tc_max) { step_mask ^= 2u; py = fsub(py, scale_exp2); }
    if (tz_corner <= tc_max) { step_mask ^= 4u; pz = fsub(pz, scale_exp2); }
    if (step_mask == 0u) {
      // All three corners are NaN (NaN direction from a zero or 555 normal):
      // nothing changes any more and the reference spins to the cap (:264).
      iter = (uint32_t)kMaxIterations + 1u;
      break;
    }
    t_min = tc_max;
    idx ^= step_mask;

    if ((idx & step_mask) != 0u) {  // POP :347-368
      uint32_t differing_bits = 0;
      if (step_mask & 1u) differing_bits |= __float_as_uint(px) ^ __float_as_uint(fadd(px, scale_exp2));
      if (step_mask & 2u) differing_bits |= __float_as_uint(py) ^ __float_as_uint(fadd(py, scale_exp2));
      if (step_mask & 4u) differing_bits |= __float_as_uint(pz) ^ __float_as_uint(fadd(pz, scale_exp2));
      scale = 31 - __clz(differing_bits);  // findMSB
      if (scale >= kMaxScale) break;       // left the cube: the loop condition fails next (:262)
      scale_exp2 = __uint_as_float((uint32_t)(scale - kMaxScale + 127) << 23);
      pidx = stk_idx[scale];
      t_max = stk_tmax[scale];
      pd = __ldg(sc.desc + pidx);
      const uint32_t shx = __float_as_uint(px) >> scale;
      const uint32_t shy = __float_as_uint(py) >> scale;
      const uint32_t shz = __float_as_uint(pz) >> scale;
      px = __uint_as_float(shx << scale);
      py = __uint_as_float(shy << scale);
      pz = __uint_as_float(shz << scale);
      idx = (shx & 1u) | ((shy & 1u) << 1) | ((shz & 1u) << 2);
      h = 0.0f;
    }
  }
  loops = iter;
  if (STATS) {
    rs->casts += 1u;
    rs->record_bytes += 7u;  // extractNode(0) :222
    rs->iters += iter > (uint32_t)kMaxIterations ? (uint32_t)kMaxIterations : iter;
  }

  if (!hit) {
    if (iter <= (uint32_t)kMaxIterations) {  // :371-377
      res.dbg = fmul(0.01f, (float)iter);
      res.dbg_init = 0;
    }
    return false;
  }

  // hit: extractChild again (:381) on the ORIGINAL bytes
  const uint32_t codes = pd.y & 0xFFFFu;
  const uint32_t ptr = __ldg(sc.refbase + pidx) + child_offset(codes, child_shift);
  const uint32_t code = (codes >> (2u * child_shift)) & 3u;
  const uint32_t value = raw_byte(sc, ptr);
  uint32_t raw16 = 0;  // Node.leafMask of the hit record
  if (code == 1u) raw16 = raw_byte(sc, ptr + 1u) | (raw_byte(sc, ptr + 2u) << 8);        // extractLeaf :103-108
  else if (code != 3u) raw16 = (raw_byte(sc, ptr + 5u) << 8) | raw_byte(sc, ptr + 6u);  // extractNode / SubdividableLeaf
  vec3 norm = mk3(0.0f, 0.0f, 0.0f);
  if (raw16 != 0u) {  // :382-388
    const int raw = (int)raw16;
    const float nx = (float)((raw % 10) - 5);
    const float ny = (float)((((raw % 100) - (raw % 10)) / 10) - 5);
    const float nz = (float)(((raw - (raw % 100)) / 100) - 5);
    norm = normalize3(mk3(nx, ny, nz));
  }
  res.pointer = ptr;
  res.t = t_min;  // :403-408
  res.value = value;
  res.iter = iter;
  res.normal = norm;
  res.scale = scale_exp2;
  res.depth = (uint32_t)(kMaxScale - scale);
  float vx = px, vy = py, vz = pz;  // :413-421
  if (d.x > 0.0f) vx = fsub(fsub(3.0f, vx), scale_exp2);
  if (d.y > 0.0f) vy = fsub(fsub(3.0f, vy), scale_exp2);
  if (d.z > 0.0f) vz = fsub(fsub(3.0f, vz), scale_exp2);
  res.voxelPos.x = fadd(vx, fmul(fmul(fmul(norm.x, scale_exp2), 2.0f), 1.74f));
  res.voxelPos.y = fadd(vy, fmul(fmul(fmul(norm.y, scale_exp2), 2.0f), 1.74f));
  res.voxelPos.z = fadd(vz, fmul(fmul(fmul(norm.z, scale_exp2), 2.0f), 1.74f));
  res.dbg = fmul(0.005f, (float)iter);  // :428
  res.dbg_init = 0;
  return true;  // :431 (scale < MAX_SCALE && t_min <= t_max both hold at either break)
}

SVO_DI void matcolor_table(uint32_t value, vec3 &mc) {  // :514-522, :578-586
  if (value == 1u) mc = mk3(0.84f, 0.86f, 0.78f);
  if (value == 2u) mc = mk3(0.57f, 0.5f, 0.31f);
  if (value == 3u) mc = mk3(0.37f, 0.43f, 0.27f);
}

struct PixelOut {
  vec3 color;
  float depth;
  uint32_t hit_id, iter;
  float primary_t;
};

SVO_DI vec3 sky(vec3 dir) {  // :449-450, :629-631
  return mk3(fsub(0.6725f, fmul(dir.y, 0.4f)), fsub(0.8784f, fmul(dir.y, 0.4f)), fsub(1.0f, fmul(dir.y, 0.25f)));
}

// trace (svotrace.comp:435-646)
template <bool FAST, bool STATS>
__device__ __forceinline__ void trace_pixel(const SceneView &sc, const FrameParams &f, float beamDist, vec3 origin,
                                            vec3 dir, float seed0, float seed1, float seed2, PixelOut &out, RayStats *rs) {
  CastRes res;
  res.value = res.pointer = res.iter = res.depth = 0u;  // uninitialised upstream; zero by contract
  res.t = 2.0f;                                          // :437
  res.scale = 0.0f;
  res.normal = mk3(0.0f, 0.0f, 0.0f);
  res.voxelPos = mk3(0.0f, 0.0f, 0.0f);
  res.dbg = 0.0f;
  res.dbg_init = 0;
  origin = mk3(fadd(origin.x, fmul(dir.x, beamDist)), fadd(origin.y, fmul(dir.y, beamDist)),
               fadd(origin.z, fmul(dir.z, beamDist)));  // :438
  const int mode = f.renderMode;
  uint32_t loops = 0;

  if (mode == 0) {  // :443-560
    vec3 accum = mk3(0.0f, 0.0f, 0.0f), mask = mk3(1.0f, 1.0f, 1.0f);
    const float is3 = fdiv(1.0f, fsqrt(3.0f));
    const vec3 sun_dir = mk3(is3, is3, is3);  // :546
    for (int i = 0; i < f.casts; i++) {
      const bool intersect = cast_ray<FAST, STATS>(sc, origin, dir, f.maxDepth, i != 0, f.coneDepth, res, loops, rs);
      if (i == 0) {
        out.iter = loops;
        out.hit_id = intersect ? res.pointer : kNoHit;
        out.primary_t = intersect ? res.t : 0.0f;
      }
      if (!intersect && i == 0) {  // :448-452
        const vec3 s = sky(dir);
        accum = mk3(fadd(accum.x, s.x), fadd(accum.y, s.y), fadd(accum.z, s.z));
        break;
      }
      const vec3 normal = res.normal;      // :476 (stale on a bounce miss, as upstream)
      const vec3 hitpoint = res.voxelPos;  // :481
      const float ra = det_rand(seed0, fmul(seed2, 0.1f));  // :486
      const float rb = det_rand(seed1, fmul(seed2, 0.02f));
      const float rnd = det_rand(fadd(seed0, ra), fadd(seed1, rb));
      const float rand1 = fmul(fmul(2.0f, 3.14159265359f), rnd);  // :487
      const vec3 w = normal;                                        // :494-497
      const vec3 axis = fabsf(w.x) > 0.1f ? mk3(0.0f, 1.0f, 0.0f) : mk3(1.0f, 0.0f, 0.0f);
      const vec3 u = normalize3(cross3(axis, w));
      const vec3 v = cross3(w, u);
      vec3 newdir;
      if (f.mirrorValue != 0 && res.value == (uint32_t)f.mirrorValue) {  // :500-504 (commented out upstream)
        const float dn = fmul(2.0f, dot3(dir, normal));
        newdir = mk3(fsub(dir.x, fmul(dn, normal.x)), fsub(dir.y, fmul(dn, normal.y)), fsub(dir.z, fmul(dn, normal.z)));
      } else {  // :506
        const float c = det_cos(rand1), s = det_sin(rand1);
        const float omr = fsub(1.0f, rnd);
        newdir = normalize3(mk3(fadd(fadd(fmul(u.x, c), fmul(v.x, s)), fmul(w.x, omr)),
                                fadd(fadd(fmul(u.y, c), fmul(v.y, s)), fmul(w.y, omr)),
                                fadd(fadd(fmul(u.z, c), fmul(v.z, s)), fmul(w.z, omr))));
      }
      origin = hitpoint;  // :508-509
      dir = newdir;
      vec3 matcolor = mk3(fsub(hitpoint.x, 1.0f), fsub(hitpoint.y, 1.0f), fsub(hitpoint.z, 1.0f));  // :511
      matcolor_table(res.value, matcolor);
      if (intersect) {  // :531-535
        out.depth = res.t;
        const float dnn = dot3(newdir, normal);
        accum = mk3(fadd(accum.x, fmul(mask.x, 0.0f)), fadd(accum.y, fmul(mask.y, 0.0f)), fadd(accum.z, fmul(mask.z, 0.0f)));
        mask = mk3(fmul(fmul(mask.x, matcolor.x), dnn), fmul(fmul(mask.y, matcolor.y), dnn), fmul(fmul(mask.z, matcolor.z), dnn));
      } else {  // :536-557
        const float diff = det_acos(dot3(dir, sun_dir));
        if (diff < 0.4f)
          accum = mk3(fadd(accum.x, fmul(mask.x, 7.0f)), fadd(accum.y, fmul(mask.y, 7.0f)), fadd(accum.z, fmul(mask.z, 7.0f)));
        accum = mk3(fadd(accum.x, fmul(mask.x, 1.0f)), fadd(accum.y, fmul(mask.y, 1.0f)), fadd(accum.z, fmul(mask.z, 1.0f)));
        out.depth = 0.0f;
        break;
      }
    }
    out.color = accum;
    return;
  }

  if (mode == 1 || mode == 2 || mode == 3) {
    const bool hit = cast_ray<FAST, STATS>(sc, origin, dir, f.maxDepth, false, f.coneDepth, res, loops, rs);
    out.iter = loops;
    out.hit_id = hit ? res.pointer : kNoHit;
    out.primary_t = hit ? res.t : 0.0f;
    if (mode == 1) {  // :561-571
      out.depth = hit ? res.t : 0.0f;
      out.color = res.dbg_init ? mk3(0.3f, 0.3f, 0.6f) : mk3(res.dbg, res.dbg, res.dbg);
      return;
    }
    if (mode == 3) {  // :633-642
      out.depth = hit ? res.t : 0.0f;
      out.color = hit ? mk3(fadd(fmul(res.normal.x, 0.5f), 0.5f), fadd(fmul(res.normal.y, 0.5f), 0.5f),
                            fadd(fmul(res.normal.z, 0.5f), 0.5f))
                      : mk3(0.0f, 0.0f, 0.0f);
      return;
    }
    if (!hit) {  // :626-632
      out.depth = 0.0f;
      out.color = sky(dir);
      return;
    }
    out.depth = res.t;  // :573-625
    vec3 matcolor = mk3(0.0f, 0.0f, 0.0f);
    matcolor_table(res.value, matcolor);
    const float sd = fdiv(0.5f, fsqrt(0.75f));  // normalize(vec3(0.5)) :587
    const vec3 sun2 = mk3(sd, sd, sd);
    float ph;
    if (res.depth >= 10u) ph = fmul(dot3(res.normal, sun2), 0.1f);  // :588-593
    else ph = fmul(dot3(mk3(0.0f, 1.0f, 0.0f), sun2), 0.1f);
    matcolor = mk3(fadd(matcolor.x, ph), fadd(matcolor.y, ph), fadd(matcolor.z, ph));
    const float base = fmul(-0.5f, fadd(res.t, beamDist));  // :595-598
    const float lambdag = det_exp(fmul(base, 2.0f));
    const float lambdab = det_exp(fmul(base, 4.0f));
    const float lambdar = det_exp(fmul(base, 1.0f));
    matcolor.x = fadd(fmul(lambdar, matcolor.x), fmul(fsub(1.0f, lambdar), 1.0f));  // :602-604
    matcolor.y = fadd(fmul(lambdag, matcolor.y), fmul(fsub(1.0f, lambdag), 1.0f));
    matcolor.z = fadd(fmul(lambdab, matcolor.z), fmul(fsub(1.0f, lambdab), 1.0f));
    const vec3 so = res.voxelPos;
    const bool shit = cast_ray<FAST, STATS>(sc, so, sun2, f.maxDepth, false, f.coneDepth, res, loops, rs);  // :607
    if (shit && res.t > fmul(res.scale, 1.73205080757f)) {
      matcolor = mk3(fsub(matcolor.x, 0.2f), fsub(matcolor.y, 0.2f), fsub(matcolor.z, 0.2f));
    } else if (res.iter > 260u) {  // :616-619
      const float pen = fdiv(fmul(0.05f, (float)res.iter), 100.0f);
      matcolor = mk3(fsub(matcolor.x, pen), fsub(matcolor.y, pen), fsub(matcolor.z, pen));
    }
    out.color = matcolor;
    return;
  }
  out.color = res.voxelPos;  // mode 4 (:643-645): uninitialised upstream, zero here
}

SVO_DI unsigned char quant8(float c) {  // imageStore to rgba8 (:726), DESIGN.md U5
  if (c != c) return 0;
  c = fminf(fmaxf(c, 0.0f), 1.0f);
  return (unsigned char)floorf(fadd(fmul(c, 255.0f), 0.5f));
}

// main (svotrace.comp:649-729) for pixel (x, y)
template <bool FAST, bool AUX, bool STATS = false>
__device__ __forceinline__ void shade_pixel(const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H,
                                            int x, int y, RayStats *rs = nullptr) {
  float beamDist = 0.0f;
  if (f.useBeam && pl.beam) beamDist = __ldg(pl.beam + (size_t)(y >> 2) * (size_t)(W >> 2) + (size_t)(x >> 2));  // :656-658
  const float fx = fdiv(fadd((float)x, 0.5f), (float)W);  // :662
  const float fy = fdiv(fadd((float)y, 0.5f), (float)H);
  vec3 dir;  // :664
  dir.x = mixf(mixf(f.l1[0], f.l2[0], fy), mixf(f.r1[0], f.r2[0], fy), fx);
  dir.y = mixf(mixf(f.l1[1], f.l2[1], fy), mixf(f.r1[1], f.r2[1], fy), fx);
  dir.z = mixf(mixf(f.l1[2], f.l2[2], fy), mixf(f.r1[2], f.r2[2], fy), fx);
  dir = normalize3(dir);  // :675
  PixelOut o;
  o.color = mk3(0.0f, 0.0f, 0.0f);
  o.depth = -1.0f;  // :672
  o.hit_id = kNoHit;
  o.iter = 0;
  o.primary_t = 0.0f;
  trace_pixel<FAST, STATS>(sc, f, beamDist, mk3(f.camPos[0], f.camPos[1], f.camPos[2]), dir, (float)x, (float)y,
                           (float)f.frameNumber, o, rs);
  if (x < 10 && y < 10)  // :696-700
    o.color = sc.first_word_zero ? mk3(1.0f, 0.0f, 0.0f) : mk3(1.0f, 1.0f, 1.0f);
  const size_t p = (size_t)y * (size_t)W + (size_t)x;
  pl.rgba8[p] = make_uchar4(quant8(o.color.x), quant8(o.color.y), quant8(o.color.z), 255);  // :726
  pl.depth[p] = o.depth;                                                                       // :727
  if (AUX) {
    pl.hit_id[p] = o.hit_id;
    pl.iter[p] = o.iter;
    pl.primary_t[p] = o.primary_t;
    pl.radiance[p] = make_float4(o.color.x, o.color.y, o.color.z, 1.0f);
  }
}

}  // namespace svo
