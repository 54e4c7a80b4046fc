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
//  * The traversal is a resumable state machine (Trav::setup / step / finish)
//    and so is the per-pixel shading (pixel_begin / pixel_after_cast /
//    pixel_store): the tile kernel runs them back to back, the persistent
//    kernel interleaves pixels so that a lane whose ray has finished is
//    refilled instead of idling until the slowest ray of its warp is done.
//  * Work that the reference repeats every iteration but that only matters
//    at certain points runs at those points: the sticky LOD-cut test where
//    t_min changes (ADVANCE), the 1500-iteration cap on the POP path (plus one
//    fix-up after the loop), the NaN-ray test once before the loop; position
//    updates are predicated FADDs (FMA pipe: the kernel is ALU-pipe bound),
//    the stack holds 8-byte (parent, t_max) entries.
//  * Casts that cannot hit anything end early: a ray outside the bounding box
//    of everything hittable (computed at upload) is a miss without walking the
//    reference's 256^3 empty cells (only where the iteration count of a miss
//    is unobservable, see Trav), and casts whose hit record trace() never
//    reads skip its decoding and the dead shading (cast_needs_attrs).
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

SVO_DI void cast_res_clear(CastRes &r) {  // uninitialised upstream; zero by contract
  r.value = r.pointer = r.iter = r.depth = 0u;
  r.t = 0.0f;
  r.scale = 0.0f;
  r.normal = mk3(0.0f, 0.0f, 0.0f);
  r.voxelPos = mk3(0.0f, 0.0f, 0.0f);
  r.dbg = 0.0f;
  r.dbg_init = 0;
}

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

SVO_DI uint32_t find_msb(uint32_t x) {  // GLSL findMSB for x != 0
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
#else  // host pass of nvcc (never called) and the g++ build of this header in tests/hostemu
  return x ? 31u - (uint32_t)__builtin_clz(x) : 0xFFFFFFFFu;
#endif
}

// a * b + c as an IMAD (FMA pipe).  Callers pass `one` -- a kernel-parameter word holding 1 that ptxas cannot fold --
// as a factor, so that an add or a move survives as a multiply-add instead of being turned back into IADD3 / MOV.
SVO_DI uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return a * b + c;
#endif
}
// if (a > b) { p += s (one rounding); bits += one * bit; }  -- one FSETP, a predicated FADD and a predicated IMAD.
// `bit` comes from the kernel parameters too (SceneView::one/two/four): with an immediate power of two ptxas emits LEA (ALU pipe).
SVO_DI void step_if_gt(float a, float b, float &p, float s, uint32_t &bits, uint32_t one, uint32_t bit) {
#ifdef __CUDA_ARCH__
  asm("{\n\t.reg .pred q;\n\tsetp.gt.f32 q, %2, %3;\n\t@q add.rn.f32 %0, %0, %4;\n\t@q mad.lo.u32 %1, %5, %6, %1;\n\t}"
      : "+f"(p), "+r"(bits)
      : "f"(a), "f"(b), "f"(s), "r"(one), "r"(bit));
#else
  if (a > b) { p = fadd(p, s); bits += one * bit; }
#endif
}

SVO_DI float fmove(float x, uint32_t one) { return __uint_as_float(imad(__float_as_uint(x), one, 0u)); }  // a register move as IMAD
// if (a <= b) { p -= s (one rounding); bits += one * bit; }
SVO_DI void step_if_le(float a, float b, float &p, float s, uint32_t &bits, uint32_t one, uint32_t bit) {
#ifdef __CUDA_ARCH__
  asm("{\n\t.reg .pred q;\n\tsetp.le.f32 q, %2, %3;\n\t@q sub.rn.f32 %0, %0, %4;\n\t@q mad.lo.u32 %1, %5, %6, %1;\n\t}"
      : "+f"(p), "+r"(bits)
      : "f"(a), "f"(b), "f"(s), "r"(one), "r"(bit));
#else
  if (a <= b) { p = fsub(p, s); bits += one * bit; }
#endif
}
// if (a > b) dst = src  -- FSETP + predicated IMAD instead of FSETP + SEL
SVO_DI void move_if_gt(float a, float b, int &dst, int src, uint32_t one) {
#ifdef __CUDA_ARCH__
  asm("{\n\t.reg .pred q;\n\tsetp.gt.f32 q, %1, %2;\n\t@q mad.lo.s32 %0, %3, %4, 0;\n\t}" : "+r"(dst) : "f"(a), "f"(b), "r"(src), "r"(one));
#else
  if (a > b) dst = src * (int)one;
#endif
}

SVO_DI float sign_glsl(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

struct RayStats {  // per-thread counters of the instrumented build (svo_render_stats)
  uint32_t casts, iters, record_bytes;
};

enum { TRAV_CONTINUE = 0, TRAV_HIT = 1, TRAV_MISS = 2 };

// State of one cast when its loop ends: enough to run the code after the loop.
struct HitState {
  uint32_t pidx;  // descriptor index of the parent of the hit child
  uint32_t meta;  // status [0:2) | child_shift [2:5) | octant_mask [5:8) | scale [8:13)
  uint32_t ipx, ipy, ipz;
  float t_min;
  uint32_t iter;
};

// after the loop (svotrace.comp:371-431).  Returns hit; `loops` = iterations run.
// `attrs` = false skips the reads of the hit record (value, packed normal) and everything derived from them: used
// for casts whose shading provably never looks at them (cast_needs_attrs).
SVO_DI bool finish_hit(const SceneView &sc, const HitState hs, CastRes &res, uint32_t &loops, bool attrs = true) {
  const uint32_t iter = hs.iter;
  loops = iter;
  if ((hs.meta & 3u) != (uint32_t)TRAV_HIT) {
    if (iter <= (uint32_t)kMaxIterations) {  // :371-377
      res.dbg = fmul(0.01f, (float)iter);
      res.dbg_init = 0;
    } else {
      res.dbg_init = 1;  // :213 survives the cap exit
    }
    return false;
  }
  const uint32_t cs = (hs.meta >> 2) & 7u, oct = (hs.meta >> 5) & 7u;
  const int scale = (int)((hs.meta >> 8) & 31u);
  const float scale_exp2 = __uint_as_float((uint32_t)(scale - kMaxScale + 127) << 23);
  if (!attrs) {  // only what trace() reads after a cast whose hit record it ignores: t, scale, iter, depth
    res.t = hs.t_min;
    res.iter = iter;
    res.scale = scale_exp2;
    res.depth = (uint32_t)(kMaxScale - scale);
    res.dbg = fmul(0.005f, (float)iter);
    res.dbg_init = 0;
    return true;
  }
  // hit: extractChild again (:381) on the ORIGINAL bytes
  const uint32_t codes = __ldg(sc.desc + hs.pidx).y & 0xFFFFu;
  const uint32_t ptr = __ldg(sc.refbase + hs.pidx) + child_offset(codes, cs);
  const uint32_t code = (codes >> (2u * cs)) & 3u;
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
  res.t = hs.t_min;  // :403-408
  res.value = value;
  res.iter = iter;
  res.normal = norm;
  res.scale = scale_exp2;
  res.depth = (uint32_t)(kMaxScale - scale);
  float vx = __uint_as_float(hs.ipx), vy = __uint_as_float(hs.ipy), vz = __uint_as_float(hs.ipz);  // :413-421
  if (oct & 1u) vx = fsub(fsub(3.0f, vx), scale_exp2);  // dir.x > 0
  if (oct & 2u) vy = fsub(fsub(3.0f, vy), scale_exp2);
  if (oct & 4u) vz = fsub(fsub(3.0f, vz), scale_exp2);
  res.voxelPos.x = fadd(vx, fmul(fmul(fmul(norm.x, scale_exp2), 2.0f), 1.74f));
  res.voxelPos.y = fadd(vy, fmul(fmul(fmul(norm.y, scale_exp2), 2.0f), 1.74f));
  res.voxelPos.z = fadd(vz, fmul(fmul(fmul(norm.z, scale_exp2), 2.0f), 1.74f));
  res.dbg = fmul(0.005f, (float)iter);  // :428
  res.dbg_init = 0;
  return true;  // :431 (scale < MAX_SCALE && t_min <= t_max both hold at either break)
}

// The parent stack (octstack, svotrace.comp:199-208) behind a small policy, so that where and how wide it is stored
// can be varied without touching the traversal: a per-thread array of 8-byte (parent index, t_max) entries in local
// memory (default); WideStack: 16-byte entries that also carry the parent's descriptor, so a POP is one load instead
// of a load followed by a dependent descriptor fetch; SmemStack: the default entries in shared memory, one column
// per thread (conflict-free), for frames whose maxDepth keeps the stack within kSmemStackLevels entries.
struct WideStack { uint4 *p; };
constexpr int kSmemStackLevels = 14, kSmemStackStride = 128;  // scales 9..22 (maxDepth <= 14); CTAs of 128 threads
struct SmemStack { uint2 *p; };                               // &smem[0][threadIdx.x]; entry of scale s at p[(s - 9) * 128]
SVO_DI void stk_store(uint2 *s, int scale, uint32_t pidx, float t_max, uint2) { s[scale] = make_uint2(pidx, __float_as_uint(t_max)); }
SVO_DI void stk_store(WideStack s, int scale, uint32_t pidx, float t_max, uint2 pd) {
  s.p[scale] = make_uint4(pidx, __float_as_uint(t_max), pd.x, pd.y);
}
SVO_DI void stk_store(SmemStack s, int scale, uint32_t pidx, float t_max, uint2) {
  s.p[(scale - (kMaxScale - kSmemStackLevels)) * kSmemStackStride] = make_uint2(pidx, __float_as_uint(t_max));
}
// returns the popped parent's descriptor; FETCH reads it from the descriptor array where the stack does not carry it
template <class FETCH>
SVO_DI uint2 stk_load(const uint2 *s, int scale, uint32_t &pidx, float &t_max, const SceneView &sc) {
  const uint2 se = s[scale];
  pidx = se.x;
  t_max = __uint_as_float(se.y);
  return FETCH::fetch(sc, pidx);
}
template <class FETCH>
SVO_DI uint2 stk_load(WideStack s, int scale, uint32_t &pidx, float &t_max, const SceneView &) {
  const uint4 se = s.p[scale];
  pidx = se.x;
  t_max = __uint_as_float(se.y);
  return make_uint2(se.z, se.w);
}
template <class FETCH>
SVO_DI uint2 stk_load(SmemStack s, int scale, uint32_t &pidx, float &t_max, const SceneView &sc) {
  const uint2 se = s.p[(scale - (kMaxScale - kSmemStackLevels)) * kSmemStackStride];
  pidx = se.x;
  t_max = __uint_as_float(se.y);
  return FETCH::fetch(sc, pidx);
}

// One intersectOctree call (svotrace.comp:211-432) as a resumable state machine.
// BOX: also use the frame's content box (SceneView::box_*): a cast whose ray never enters the box, or has left
// it, cannot hit anything and is ended as a miss at once.  Only the iteration count of a MISSING cast differs
// from the reference, so callers enable it only where that count is unobservable (not render mode 1, no
// validation planes, no ray-stream API).
// TOP: the first sc.ntop descriptors (the upper octree levels: the array is breadth-first) are read from a
// shared-memory copy (sc.top) instead of global memory.
// BAL: pipe-balanced loop body.  ncu puts the ALU pipe (LOP3 / SHF / SEL / FSETP / FMNMX / IADD3 / MOV) at 75 % busy with
// math-pipe throttle among the top stall reasons while the FMA pipe idles at 20 %; BAL moves the integer adds, moves
// and select chains of the loop onto the FMA pipe as IMADs (imad() below).  Same values, different instructions.
// BAL is a mask so that the parts can be measured separately: 1 = child-index / step-mask assembly as predicated FADD + IMAD
// (step_if_gt / step_if_le / move_if_gt), 2 = the h / t_max / t_min moves as IMADs, 4 = pidx and scale-1 in PUSH as IMADs,
// 8 = the POP path's shifts, 2^(scale-23) and child index as IMADs.  15 = kernel variant 13.
// STATS: 1 = the oracle's counters (what the REFERENCE would have fetched: a NaN ray counts its 1500 spins), 2 = what this
// kernel executes (iterations actually run, with the content box and the early exits).
template <bool FAST, int STATS = 0, bool BOX = false, bool TOP = false, int BAL = 0>
struct Trav {
  typedef Ops<FAST> M;
  static __device__ __forceinline__ uint2 fetch(const SceneView &sc, uint32_t i) {
    if (TOP && i < sc.ntop) return sc.top[i];
    return __ldg(sc.desc + i);
  }
  float tb_out;             // BOX: ray parameter at which the ray leaves the content box
  float cx, cy, cz;         // t*_coef
  float bx, by, bz;         // t*_bias
  float t_min, t_max, h;
  float px, py, pz;         // pos
  float scale_exp2;
  int scale;
  int stop_scale;           // MAX_SCALE - maxDepth
  int cone_stop;            // what stop_scale becomes once t_min > 0.05 (== stop_scale when !coneTrace)
  uint32_t idx, oct, pidx;
  float iter;               // loop iterations so far, counted in binary32 (exact: the cap is 1500) so that the
                            // increment runs on the FMA pipe -- the loop is bound by the ALU pipe
  uint2 pd;

  // t_floor > 0: a proven lower bound on the cast's hit distance (conservative beam pre-pass): the walk starts at the cell
  // that contains the ray at that parameter instead of at the cube's entry.  Same hit, fewer iterations.
  __device__ __forceinline__ void setup(const SceneView &sc, const vec3 o, vec3 d, int maxDepth, bool coneTrace,
                                        int coneDepth, RayStats *rs, float t_floor = 0.0f) {
    const float kEps = 3.552713678800501e-15f;                // :31
    if (fabsf(d.x) < kEps) d.x = fmul(kEps, sign_glsl(d.x));  // :226-228
    if (fabsf(d.y) < kEps) d.y = fmul(kEps, sign_glsl(d.y));
    if (fabsf(d.z) < kEps) d.z = fmul(kEps, sign_glsl(d.z));
    cx = fdiv(1.0f, -fabsf(d.x));  // :230-232
    cy = fdiv(1.0f, -fabsf(d.y));
    cz = fdiv(1.0f, -fabsf(d.z));
    bx = M::mul(cx, o.x);  // :234-236
    by = M::mul(cy, o.y);
    bz = M::mul(cz, o.z);
    oct = 0;  // :238-241
    if (d.x > 0.0f) { oct ^= 1u; bx = M::msub(3.0f, cx, bx); }
    if (d.y > 0.0f) { oct ^= 2u; by = M::msub(3.0f, cy, by); }
    if (d.z > 0.0f) { oct ^= 4u; bz = M::msub(3.0f, cz, bz); }
    t_min = fmaxf(fmaxf(M::msub(2.0f, cx, bx), M::msub(2.0f, cy, by)), M::msub(2.0f, cz, bz));  // :243
    t_max = fminf(fminf(M::sub(cx, bx), M::sub(cy, by)), M::sub(cz, bz));                       // :244
    t_min = fmaxf(t_min, 0.0f);                                                                  // :245
    if (t_floor > 0.0f) t_min = fmaxf(t_min, t_floor);
    h = t_max;                                                                                   // :247
    idx = 0;
    px = py = pz = 1.0f;
    scale = kMaxScale - 1;
    scale_exp2 = 0.5f;
    if (M::msub(1.5f, cx, bx) > t_min) { idx ^= 1u; px = 1.5f; }  // :255-257
    if (M::msub(1.5f, cy, by) > t_min) { idx ^= 2u; py = 1.5f; }
    if (M::msub(1.5f, cz, bz) > t_min) { idx ^= 4u; pz = 1.5f; }
    pidx = 0;               // parent = root (:222)
    pd = fetch(sc, 0u);
    iter = 0.0f;
    stop_scale = kMaxScale - maxDepth;
    cone_stop = coneTrace ? kMaxScale - coneDepth : stop_scale;
    // the sticky LOD cut (:275-277) is tested at the top of every iteration upstream; t_min only changes here
    // and in ADVANCE, so testing it at those two places is the same thing
    if (t_min > 0.05f) stop_scale = cone_stop;
    if (STATS) { rs->casts += 1u; rs->record_bytes += 7u; }  // extractNode(0) :222
    if (BOX) {
      // slab test in the mirrored coordinates of the traversal (the ray moves towards smaller coordinates on every
      // axis): it enters the box through the high faces and leaves through the low ones.  The box is padded by
      // 2^-9; rays with a direction component below 2^-8 or an origin coordinate beyond 8 are exempt (their t
      // arithmetic is too ill-conditioned to bound how far the reference's cell walk strays from the true ray: with
      // |coef| <= 256 and |origin| <= 8 every t is exact to ~2.4e-4, an order of magnitude inside the padding).
      const float hx = (oct & 1u) ? 3.0f - sc.box_lo[0] : sc.box_hi[0], lx = (oct & 1u) ? 3.0f - sc.box_hi[0] : sc.box_lo[0];
      const float hy = (oct & 2u) ? 3.0f - sc.box_lo[1] : sc.box_hi[1], ly = (oct & 2u) ? 3.0f - sc.box_hi[1] : sc.box_lo[1];
      const float hz = (oct & 4u) ? 3.0f - sc.box_lo[2] : sc.box_hi[2], lz = (oct & 4u) ? 3.0f - sc.box_hi[2] : sc.box_lo[2];
      const float tb_in = fmaxf(fmaxf(hx * cx - bx, hy * cy - by), fmaxf(hz * cz - bz, 0.0f));
      tb_out = fminf(fminf(lx * cx - bx, ly * cy - by), lz * cz - bz);
      const float dmin = fminf(fminf(fabsf(d.x), fabsf(d.y)), fabsf(d.z));
      const float omax = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
      // exempt: ill-conditioned directions, far-away origins (|t_bias| grows with |o| and so does the rounding
      // error of every t), NaNs
      if (!(dmin >= 0.00390625f) || !(omax <= 8.0f)) tb_out = __uint_as_float(0x7f800000u);
      else if (!(tb_in <= tb_out)) tb_out = -1.0f;                         // never inside the box: ends at the first POP/ADVANCE test
    }
  }
  // The state setup() leaves behind, in ten words, and back: lets one thread set a cast up and another one trace it
  // (kernel variant 16).  Everything else setup() writes is a constant or follows from these and the frame.
  __device__ __forceinline__ void save_setup(uint32_t w[10]) const {
    w[0] = __float_as_uint(cx); w[1] = __float_as_uint(cy); w[2] = __float_as_uint(cz);
    w[3] = __float_as_uint(bx); w[4] = __float_as_uint(by); w[5] = __float_as_uint(bz);
    w[6] = __float_as_uint(t_min); w[7] = __float_as_uint(t_max); w[8] = BOX ? __float_as_uint(tb_out) : 0u;
    w[9] = oct | (idx << 3);
  }
  __device__ __forceinline__ void restore_setup(const SceneView &sc, const uint32_t w[10], int maxDepth, bool coneTrace, int coneDepth) {
    cx = __uint_as_float(w[0]); cy = __uint_as_float(w[1]); cz = __uint_as_float(w[2]);
    bx = __uint_as_float(w[3]); by = __uint_as_float(w[4]); bz = __uint_as_float(w[5]);
    t_min = __uint_as_float(w[6]); t_max = __uint_as_float(w[7]);
    if (BOX) tb_out = __uint_as_float(w[8]);
    oct = w[9] & 7u;
    idx = (w[9] >> 3) & 7u;
    h = t_max;
    px = (idx & 1u) ? 1.5f : 1.0f;
    py = (idx & 2u) ? 1.5f : 1.0f;
    pz = (idx & 4u) ? 1.5f : 1.0f;
    scale = kMaxScale - 1;
    scale_exp2 = 0.5f;
    pidx = 0;
    pd = fetch(sc, 0u);
    iter = 0.0f;
    stop_scale = kMaxScale - maxDepth;
    cone_stop = coneTrace ? kMaxScale - coneDepth : stop_scale;
    if (t_min > 0.05f) stop_scale = cone_stop;
  }
  // true if the cast can end now: nothing can be hit any more
  __device__ __forceinline__ bool outside_box() const { return BOX && t_min > tb_out; }
  // All three corner times NaN (NaN direction from a zero or 555 normal, NaN origin, or a zero direction): that does
  // not depend on pos, so it holds for the whole cast; no child ever passes t_min <= t_max, ADVANCE never steps,
  // and the reference spins to the iteration cap (:264).  Callers of run() test it once instead of every iteration.
  __device__ __forceinline__ bool nan_ray(RayStats *rs) {
    const float a = M::msub(px, cx, bx), b = M::msub(py, cy, by), c = M::msub(pz, cz, bz);
    if (!((a != a) && (b != b) && (c != c))) return false;
    if (STATS == 1) {  // the reference fetches the same child record 1500 times
      const uint32_t code = (pd.y >> (2u * (idx ^ oct))) & 3u;
      rs->iters += (uint32_t)kMaxIterations;
      rs->record_bytes += (uint32_t)kMaxIterations * (code == 1u ? 3u : (code == 3u ? 1u : 7u));
    }
    iter = (float)(kMaxIterations + 1);
    return true;
  }

  // The body of one loop iteration (:262-369), shared by run() and step().  `EXIT(status)` leaves the loop,
  // `NEXT` starts the next iteration.  The iteration cap (:264-266) is enforced where it is cheap -- on the
  // POP path, plus once after the loop (cap_fixup) -- instead of on every iteration: a cast that ends with
  // iter > 1500 is exactly a cast the reference capped, and a cast cannot run long without a POP.
#define SVO_TRAV_BODY(EXIT, NEXT, NANCHECK)                                                                                  \
  iter = fadd(iter, 1.0f);                                                                                           \
  if (STATS && iter <= (float)kMaxIterations) rs->iters += 1u;                                                       \
  const float tx_corner = M::msub(px, cx, bx); /* :280-283 */                                                        \
  const float ty_corner = M::msub(py, cy, by);                                                                       \
  const float tz_corner = M::msub(pz, cz, bz);                                                                       \
  const float tc_max = fminf(fminf(tx_corner, ty_corner), tz_corner);                                                \
  const uint32_t cs = idx ^ oct; /* child_shift :286 */                                                              \
  if (STATS && iter <= (float)kMaxIterations) { /* size of the child record fetched here (extractChild :294) */     \
    const uint32_t code = (pd.y >> (2u * cs)) & 3u;                                                                  \
    rs->record_bytes += code == 1u ? 3u : (code == 3u ? 1u : 7u);                                                    \
  }                                                                                                                  \
  const uint32_t m = pd.y >> cs;                                                                                     \
  /* child.value != 0 (:295) is bit 16+child of the parent's descriptor */                                           \
  if ((m & 0x10000u) != 0u && t_min <= t_max) {                                                                      \
    if (scale == stop_scale) { EXIT(TRAV_HIT); } /* MAX_SCALE - scale == maxDepth (:300-302) */                      \
    const float tv_max = fminf(t_max, tc_max);   /* :304 */                                                          \
    if (t_min <= tv_max) {                       /* :310 */                                                          \
      if ((m & 0x01000000u) == 0u) { EXIT(TRAV_HIT); } /* child.cp == 0 (:311-313) */                                \
      if (tc_max < h) { /* PUSH :316-319 */                                                                          \
        stk_store(stk, scale, pidx, t_max, pd);                                                                      \
      }                                                                                                              \
      h = (BAL & 2) ? fmove(tc_max, sc.one) : tc_max;                                                                      \
      /* descriptors of the interior siblings below: bits [24, 24+cs) -- the funnel shift leaves exactly those */    \
      pidx = (BAL & 4) ? imad((uint32_t)__popc(__funnelshift_r(0u, pd.y >> 24, cs)), sc.one, pd.x)                         \
                 : pd.x + __popc(__funnelshift_r(0u, pd.y >> 24, cs));                                               \
      pd = fetch(sc, pidx);                                /* parent = child (:322) */                               \
      const float half = M::mul(scale_exp2, 0.5f);                                                                   \
      const float tx_center = M::madd(half, cx, tx_corner); /* :306-308 */                                           \
      const float ty_center = M::madd(half, cy, ty_corner);                                                          \
      const float tz_center = M::madd(half, cz, tz_corner);                                                          \
      scale_exp2 = half;                                                                                             \
      if (BAL & 4) scale = (int)imad((uint32_t)scale, sc.one, 0xFFFFFFFFu); else --scale;                            \
      if (BAL & 1) {                                                                                                 \
        idx = sc.zero; /* a zero the compiler cannot see: keeps the three steps below predicated IMADs */          \
        step_if_gt(tx_center, t_min, px, scale_exp2, idx, sc.one, sc.one); /* :328-330, exact adds */                \
        step_if_gt(ty_center, t_min, py, scale_exp2, idx, sc.one, sc.two);                                           \
        step_if_gt(tz_center, t_min, pz, scale_exp2, idx, sc.one, sc.four);                                          \
      } else {                                                                                                       \
        const bool gx = tx_center > t_min, gy = ty_center > t_min, gz = tz_center > t_min; /* :328-330 */            \
        px = gx ? fadd(px, scale_exp2) : px; /* exact adds */                                                        \
        py = gy ? fadd(py, scale_exp2) : py;                                                                         \
        pz = gz ? fadd(pz, scale_exp2) : pz;                                                                         \
        idx = (gx ? 1u : 0u) | (gy ? 2u : 0u) | (gz ? 4u : 0u);                                                      \
      }                                                                                                              \
      t_max = (BAL & 2) ? fmove(tv_max, sc.one) : tv_max;                                                                  \
      NEXT;                                                                                                          \
    }                                                                                                                \
  }                                                                                                                  \
  /* ADVANCE :337-344 */                                                                                             \
  const bool sx = tx_corner <= tc_max, sy = ty_corner <= tc_max, sz = tz_corner <= tc_max;                           \
  if (NANCHECK && !(sx || sy || sz)) {                                                                               \
    /* all three corners are NaN (NaN direction from a zero or 555 normal): nothing changes any more and the   */    \
    /* reference spins to the cap (:264).  run() callers test this once, before the loop (nan_ray()).           */    \
    if (STATS == 1 && iter <= (float)kMaxIterations) {                                                               \
      const uint32_t code = (pd.y >> (2u * cs)) & 3u, left = (uint32_t)kMaxIterations - (uint32_t)iter;              \
      rs->iters += left;                                                                                             \
      rs->record_bytes += left * (code == 1u ? 3u : (code == 3u ? 1u : 7u));                                         \
    }                                                                                                                \
    iter = (float)(kMaxIterations + 1);                                                                              \
    EXIT(TRAV_MISS);                                                                                                 \
  }                                                                                                                  \
  uint32_t step_mask;                                                                                                \
  if (BAL & 1) {                                                                                                     \
    step_mask = sc.zero;                                                                                             \
    step_if_le(tx_corner, tc_max, px, scale_exp2, step_mask, sc.one, sc.one);                                        \
    step_if_le(ty_corner, tc_max, py, scale_exp2, step_mask, sc.one, sc.two);                                        \
    step_if_le(tz_corner, tc_max, pz, scale_exp2, step_mask, sc.one, sc.four);                                       \
    t_min = (BAL & 2) ? fmove(tc_max, sc.one) : tc_max;                                                              \
    move_if_gt(t_min, 0.05f, stop_scale, cone_stop, sc.one); /* :275-277 */                                          \
  } else {                                                                                                           \
    px = sx ? fsub(px, scale_exp2) : px;                                                                             \
    py = sy ? fsub(py, scale_exp2) : py;                                                                             \
    pz = sz ? fsub(pz, scale_exp2) : pz;                                                                             \
    step_mask = (sx ? 1u : 0u) | (sy ? 2u : 0u) | (sz ? 4u : 0u);                                                    \
    t_min = tc_max;                                                                                                  \
    if (t_min > 0.05f) stop_scale = cone_stop; /* :275-277 */                                                        \
  }                                                                                                                  \
  idx ^= step_mask;                                                                                                  \
  if ((idx & step_mask) != 0u) { /* POP :347-368 */                                                                  \
    /* The iteration cap (:264-266) is tested here only: every run of ADVANCEs ends in a POP after at most 3   */    \
    /* steps, so the test is at most ~26 iterations late, and cap_fixup() turns any cast that ends with        */    \
    /* iter > 1500 into exactly what the reference returns at iteration 1501.  (iter == 1500 must go on: this   */    \
    /* POP may leave the cube, and then the reference ends the cast as an ordinary miss at iteration 1500.)     */    \
    /* ... and so is "the ray has left the content box for good" (BOX).                                         */    \
    if (iter > (float)kMaxIterations || (BOX && t_min > tb_out)) {                                                   \
      EXIT(TRAV_MISS); /* cap_fixup() normalises iter */                                                             \
    }                                                                                                                \
    uint32_t differing_bits = 0;                                                                                     \
    if (sx) differing_bits |= __float_as_uint(px) ^ __float_as_uint(fadd(px, scale_exp2));                           \
    if (sy) differing_bits |= __float_as_uint(py) ^ __float_as_uint(fadd(py, scale_exp2));                           \
    if (sz) differing_bits |= __float_as_uint(pz) ^ __float_as_uint(fadd(pz, scale_exp2));                           \
    scale = (int)find_msb(differing_bits); /* findMSB */                                                             \
    if (scale >= kMaxScale) { EXIT(TRAV_MISS); } /* left the cube: the loop condition fails (:262) */                \
    scale_exp2 = (BAL & 8) ? __uint_as_float(imad((uint32_t)scale, sc.exp_unit, (uint32_t)(127 - kMaxScale) << 23))        \
                     : __uint_as_float((uint32_t)(scale - kMaxScale + 127) << 23);                                   \
    pd = stk_load<Trav>(stk, scale, pidx, t_max, sc);                                                                \
    const uint32_t shx = __float_as_uint(px) >> scale;                                                               \
    const uint32_t shy = __float_as_uint(py) >> scale;                                                               \
    const uint32_t shz = __float_as_uint(pz) >> scale;                                                               \
    if (BAL & 8) { /* x << scale as a multiplication by 2^scale; the child index assembled by IMADs */                  \
      const uint32_t pow2 = sc.one << scale;                                                                         \
      px = __uint_as_float(imad(shx, pow2, 0u));                                                                     \
      py = __uint_as_float(imad(shy, pow2, 0u));                                                                     \
      pz = __uint_as_float(imad(shz, pow2, 0u));                                                                     \
      idx = imad(shz & 1u, sc.four, imad(shy & 1u, sc.two, shx & 1u));                                               \
    } else {                                                                                                         \
      px = __uint_as_float(shx << scale);                                                                            \
      py = __uint_as_float(shy << scale);                                                                            \
      pz = __uint_as_float(shz << scale);                                                                            \
      idx = (shx & 1u) | ((shy & 1u) << 1) | ((shz & 1u) << 2);                                                      \
    }                                                                                                                \
    h = 0.0f;                                                                                                        \
  }

  // A hit found after iteration 1500 (only possible through a run of PUSHes right after the last ADVANCE
  // check) is a cast the reference abandoned at iteration 1501 (:264-266).
  __device__ __forceinline__ int cap_fixup(int status) {
    if (iter > (float)kMaxIterations) {
      iter = (float)(kMaxIterations + 1);
      return TRAV_MISS;
    }
    return status;
  }

  // the whole loop (:262-369); returns TRAV_HIT or TRAV_MISS
  template <class Stk>
  __device__ __forceinline__ int run(const SceneView &sc, Stk stk, RayStats *rs) {
#define SVO_EXIT(s) { if ((s) == TRAV_HIT) goto hit; else goto miss; }
    for (;;) {
      SVO_TRAV_BODY(SVO_EXIT, continue, false)
    }
#undef SVO_EXIT
  hit:
    return cap_fixup(TRAV_HIT);
  miss:
    return cap_fixup(TRAV_MISS);
  }

  // one iteration; returns TRAV_CONTINUE until the cast is over
  template <class Stk>
  __device__ __forceinline__ int step(const SceneView &sc, Stk stk, RayStats *rs) {
#define SVO_EXIT(s) return cap_fixup(s)
    SVO_TRAV_BODY(SVO_EXIT, return TRAV_CONTINUE, true)
#undef SVO_EXIT
    return TRAV_CONTINUE;
  }

  // what the code after the loop needs (:371-431), small enough to park in memory between kernels
  __device__ __forceinline__ HitState export_hit(int status) const {
    HitState hs;
    hs.pidx = pidx;
    hs.meta = (uint32_t)status | ((idx ^ oct) << 2) | (oct << 5) | ((uint32_t)scale << 8);
    hs.ipx = __float_as_uint(px);
    hs.ipy = __float_as_uint(py);
    hs.ipz = __float_as_uint(pz);
    hs.t_min = t_min;
    hs.iter = (uint32_t)iter;
    return hs;
  }
  __device__ __forceinline__ bool finish(const SceneView &sc, int status, CastRes &res, uint32_t &loops, bool attrs = true) const {
    return finish_hit(sc, export_hit(status), res, loops, attrs);
  }
};

// intersectOctree run to completion on the caller's stack.
template <bool FAST, int STATS, bool BOX, bool TOP, int BAL = 0, class Stk>
__device__ __forceinline__ bool cast_ray_on(Stk stk, const SceneView &sc, const vec3 o, const vec3 d, int maxDepth, bool coneTrace,
                                            int coneDepth, CastRes &res, uint32_t &loops, RayStats *rs, bool attrs, float t_floor = 0.0f) {
  Trav<FAST, STATS, BOX, TOP, BAL> T;
  T.setup(sc, o, d, maxDepth, coneTrace, coneDepth, rs, t_floor);
  if (T.outside_box() || T.nan_ray(rs)) return T.finish(sc, TRAV_MISS, res, loops, attrs);  // no iteration can change anything
  return T.finish(sc, T.run(sc, stk, rs), res, loops, attrs);
}
// ... on the default stack
template <bool FAST, int STATS = 0, bool BOX = false, bool TOP = false>
__device__ __forceinline__ bool cast_ray(const SceneView &sc, const vec3 o, const vec3 d, int maxDepth, bool coneTrace,
                                         int coneDepth, CastRes &res, uint32_t &loops, RayStats *rs = nullptr, bool attrs = true, float t_floor = 0.0f) {
  uint2 stk[kMaxScale + 1];  // octstack (:199-202): (parent index, t_max) per scale
  return cast_ray_on<FAST, STATS, BOX, TOP>(stk, sc, o, d, maxDepth, coneTrace, coneDepth, res, loops, rs, attrs, t_floor);
}

SVO_DI void matcolor_table(uint32_t value, vec3 &mc) {  // :514-522, :578-586
  if (value == 1u) mc = mk3(0.84f, 0.86f, 0.78f);
  if (value == 2u) mc = mk3(0.57f, 0.5f, 0.31f);
  if (value == 3u) mc = mk3(0.37f, 0.43f, 0.27f);
}

SVO_DI vec3 sky(vec3 dir) {  // :449-450, :629-631
  return mk3(fsub(0.6725f, fmul(dir.y, 0.4f)), fsub(0.8784f, fmul(dir.y, 0.4f)), fsub(1.0f, fmul(dir.y, 0.25f)));
}

// Everything one shader invocation keeps between its intersectOctree calls
// (main :649-729 + trace :435-646), so that it can be suspended at a cast.
struct Pixel {
  int x, y;
  vec3 origin, dir;  // ray of the cast about to run / running
  bool cone;         // its coneTrace argument
  int cast_i;        // which cast of the pixel this is
  vec3 acc;          // mode 0: accum_color;  mode 2: matcolor carried across the shadow cast
  vec3 mask;         // mode 0
  CastRes res;       // `res`, stale fields and all
  vec3 color;        // finalcolor
  float depth, beamDist;
  float t_floor;     // conservative beam pre-pass (svo_frame.flags bit 1): lower bound on the primary hit distance; +inf = the block misses
  uint32_t hit_id, iter;
  float primary_t;
};

// main() up to the first cast.  Returns true if the pixel wants a cast (ray in P.origin/dir/cone).
SVO_DI bool pixel_begin(const FrameParams &f, const Planes &pl, int W, int H, int x, int y, Pixel &P) {
  P.x = x;
  P.y = y;
  P.beamDist = 0.0f;
  P.t_floor = 0.0f;
  // (not in render mode 2: its penumbra term reads the PRIMARY cast's stale iteration count when the shadow ray misses,
  // svotrace.comp:616-619, so the count is observable there; mode 1 shows it outright)
  if ((f.flags & 2) && (f.renderMode == 0 || f.renderMode == 3) && pl.beam && (x >> 2) < (W >> 2) && (y >> 2) < (H >> 2))
    P.t_floor = __ldg(pl.beam + (size_t)(y >> 2) * (size_t)(W >> 2) + (size_t)(x >> 2));
  // :656-658.  The beam image has (W/4) x (H/4) texels (Main.java:82-83); imageLoad outside an image returns 0
  if (f.useBeam && pl.beam && (x >> 2) < (W >> 2) && (y >> 2) < (H >> 2))
    P.beamDist = __ldg(pl.beam + (size_t)(y >> 2) * (size_t)(W >> 2) + (size_t)(x >> 2));
  const float fx = fdiv(fadd((float)x, 0.5f), (float)W);  // :662
  const float fy = fdiv(fadd((float)y, 0.5f), (float)H);
  vec3 dir;  // :664
  dir.x = mixf(mixf(f.l1[0], f.l2[0], fy), mixf(f.r1[0], f.r2[0], fy), fx);
  dir.y = mixf(mixf(f.l1[1], f.l2[1], fy), mixf(f.r1[1], f.r2[1], fy), fx);
  dir.z = mixf(mixf(f.l1[2], f.l2[2], fy), mixf(f.r1[2], f.r2[2], fy), fx);
  dir = normalize3(dir);  // :675
  P.dir = dir;
  P.origin = mk3(fadd(f.camPos[0], fmul(dir.x, P.beamDist)), fadd(f.camPos[1], fmul(dir.y, P.beamDist)),
                 fadd(f.camPos[2], fmul(dir.z, P.beamDist)));  // :438
  P.color = mk3(0.0f, 0.0f, 0.0f);
  P.depth = -1.0f;  // :672
  P.hit_id = kNoHit;
  P.iter = 0;
  P.primary_t = 0.0f;
  cast_res_clear(P.res);
  P.res.t = 2.0f;  // :437
  P.acc = mk3(0.0f, 0.0f, 0.0f);
  P.mask = mk3(1.0f, 1.0f, 1.0f);
  P.cone = false;
  P.cast_i = 0;
  const int mode = f.renderMode;
  if (mode == 0) return f.casts > 0;         // :443-444
  if (mode >= 1 && mode <= 3) return true;   // :562, :573, :634
  P.color = P.res.voxelPos;                  // mode 4 (:643-645): uninitialised upstream, zero here
  return false;
}

// The code of trace() that follows cast number P.cast_i.  Returns true if another cast is wanted.
SVO_DI bool pixel_after_cast(const FrameParams &f, Pixel &P, bool intersect, uint32_t loops) {
  CastRes &res = P.res;
  const int mode = f.renderMode;
  if (P.cast_i == 0) {
    P.iter = loops;
    P.hit_id = intersect ? res.pointer : kNoHit;
    P.primary_t = intersect ? res.t : 0.0f;
  }
  if (mode == 0) {  // loop body :445-558
    const int i = P.cast_i;
    if (!intersect && i == 0) {  // :448-452
      const vec3 s = sky(P.dir);
      P.color = mk3(fadd(P.acc.x, s.x), fadd(P.acc.y, s.y), fadd(P.acc.z, s.z));
      return false;
    }
    const vec3 normal = res.normal;      // :476 (stale on a bounce miss, as upstream)
    const vec3 hitpoint = res.voxelPos;  // :481
    const float seed0 = (float)P.x, seed1 = (float)P.y, seed2 = (float)f.frameNumber;  // :677-681
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
      const float dn = fmul(2.0f, dot3(P.dir, normal));
      newdir = mk3(fsub(P.dir.x, fmul(dn, normal.x)), fsub(P.dir.y, fmul(dn, normal.y)), fsub(P.dir.z, fmul(dn, normal.z)));
    } else {  // :506
      const float c = det_cos(rand1), s = det_sin(rand1);
      const float omr = fsub(1.0f, rnd);
      newdir = normalize3(mk3(fadd(fadd(fmul(u.x, c), fmul(v.x, s)), fmul(w.x, omr)),
                              fadd(fadd(fmul(u.y, c), fmul(v.y, s)), fmul(w.y, omr)),
                              fadd(fadd(fmul(u.z, c), fmul(v.z, s)), fmul(w.z, omr))));
    }
    P.origin = hitpoint;  // :508-509
    P.dir = newdir;
    vec3 matcolor = mk3(fsub(hitpoint.x, 1.0f), fsub(hitpoint.y, 1.0f), fsub(hitpoint.z, 1.0f));  // :511
    matcolor_table(res.value, matcolor);
    if (intersect) {  // :531-535
      P.depth = res.t;
      const float dnn = dot3(newdir, normal);
      P.acc = mk3(fadd(P.acc.x, fmul(P.mask.x, 0.0f)), fadd(P.acc.y, fmul(P.mask.y, 0.0f)), fadd(P.acc.z, fmul(P.mask.z, 0.0f)));
      P.mask = mk3(fmul(fmul(P.mask.x, matcolor.x), dnn), fmul(fmul(P.mask.y, matcolor.y), dnn), fmul(fmul(P.mask.z, matcolor.z), dnn));
    } else {  // :536-557
      const float is3 = fdiv(1.0f, fsqrt(3.0f));
      const float diff = det_acos(dot3(P.dir, mk3(is3, is3, is3)));  // :546-547
      if (diff < 0.4f)
        P.acc = mk3(fadd(P.acc.x, fmul(P.mask.x, 7.0f)), fadd(P.acc.y, fmul(P.mask.y, 7.0f)), fadd(P.acc.z, fmul(P.mask.z, 7.0f)));
      P.acc = mk3(fadd(P.acc.x, fmul(P.mask.x, 1.0f)), fadd(P.acc.y, fmul(P.mask.y, 1.0f)), fadd(P.acc.z, fmul(P.mask.z, 1.0f)));
      P.depth = 0.0f;
      P.color = P.acc;
      return false;
    }
    P.color = P.acc;
    P.cast_i = i + 1;
    P.cone = true;  // :445-446
    return P.cast_i < f.casts;
  }
  if (mode == 1) {  // :561-571
    P.depth = intersect ? res.t : 0.0f;
    P.color = res.dbg_init ? mk3(0.3f, 0.3f, 0.6f) : mk3(res.dbg, res.dbg, res.dbg);
    return false;
  }
  if (mode == 3) {  // :633-642
    P.depth = intersect ? res.t : 0.0f;
    P.color = intersect ? mk3(fadd(fmul(res.normal.x, 0.5f), 0.5f), fadd(fmul(res.normal.y, 0.5f), 0.5f),
                              fadd(fmul(res.normal.z, 0.5f), 0.5f))
                        : mk3(0.0f, 0.0f, 0.0f);
    return false;
  }
  // mode 2 :572-632
  const float sd = fdiv(0.5f, fsqrt(0.75f));  // normalize(vec3(0.5)) :587
  if (P.cast_i == 0) {
    if (!intersect) {  // :626-632
      P.depth = 0.0f;
      P.color = sky(P.dir);
      return false;
    }
    P.depth = res.t;  // :573-604
    vec3 matcolor = mk3(0.0f, 0.0f, 0.0f);
    matcolor_table(res.value, matcolor);
    const vec3 sun2 = mk3(sd, sd, sd);
    float ph;
    if (res.depth >= 10u) ph = fmul(dot3(res.normal, sun2), 0.1f);  // :588-593
    else ph = fmul(dot3(mk3(0.0f, 1.0f, 0.0f), sun2), 0.1f);
    matcolor = mk3(fadd(matcolor.x, ph), fadd(matcolor.y, ph), fadd(matcolor.z, ph));
    const float base = fmul(-0.5f, fadd(res.t, P.beamDist));  // :595-598
    const float lambdag = det_exp(fmul(base, 2.0f));
    const float lambdab = det_exp(fmul(base, 4.0f));
    const float lambdar = det_exp(fmul(base, 1.0f));
    matcolor.x = fadd(fmul(lambdar, matcolor.x), fmul(fsub(1.0f, lambdar), 1.0f));  // :602-604
    matcolor.y = fadd(fmul(lambdag, matcolor.y), fmul(fsub(1.0f, lambdag), 1.0f));
    matcolor.z = fadd(fmul(lambdab, matcolor.z), fmul(fsub(1.0f, lambdab), 1.0f));
    P.acc = matcolor;
    P.origin = res.voxelPos;  // shadow ray :607
    P.dir = sun2;
    P.cone = false;
    P.cast_i = 1;
    return true;
  }
  vec3 matcolor = P.acc;
  if (intersect && res.t > fmul(res.scale, 1.73205080757f)) {  // :607-615
    matcolor = mk3(fsub(matcolor.x, 0.2f), fsub(matcolor.y, 0.2f), fsub(matcolor.z, 0.2f));
  } else if (res.iter > 260u) {  // :616-619
    const float pen = fdiv(fmul(0.05f, (float)res.iter), 100.0f);
    matcolor = mk3(fsub(matcolor.x, pen), fsub(matcolor.y, pen), fsub(matcolor.z, pen));
  }
  P.color = matcolor;
  return false;
}

// Does trace() read the hit record (value / normal / voxelPos) of the cast that is about to run?  Not for the
// LAST cast of the mode-0 loop -- a hit there only sets depth and multiplies a mask nobody reads again (:531-535;
// accum is returned as is) -- and not for the shadow ray of mode 2 (:607-619 reads t, scale and iter only).
SVO_DI bool cast_needs_attrs(const FrameParams &f, const Pixel &P) {
  if (f.renderMode == 0) return !(P.cast_i > 0 && P.cast_i + 1 >= f.casts);
  if (f.renderMode == 2) return P.cast_i == 0;
  return true;
}
// The code of trace() after the last cast of the mode-0 loop when that cast HIT, with the dead parts removed:
// depth = res.t; accum += mask * matemi (matemi = 0, kept because it propagates a NaN/inf mask); return accum.
SVO_DI void pixel_after_last_hit(Pixel &P) {
  P.depth = P.res.t;
  P.acc = mk3(fadd(P.acc.x, fmul(P.mask.x, 0.0f)), fadd(P.acc.y, fmul(P.mask.y, 0.0f)), fadd(P.acc.z, fmul(P.mask.z, 0.0f)));
  P.color = P.acc;
}
// finish one cast of pixel P (status known): returns true if the pixel wants another cast
SVO_DI bool pixel_finish_cast(const SceneView &sc, const FrameParams &f, Pixel &P, const HitState hs) {
  const bool attrs = cast_needs_attrs(f, P);
  uint32_t loops;
  const bool hit = finish_hit(sc, hs, P.res, loops, attrs);
  if (!attrs && hit && f.renderMode == 0) {
    pixel_after_last_hit(P);
    return false;
  }
  return pixel_after_cast(f, P, hit, loops);
}

SVO_DI unsigned char quant8(float c) {  // imageStore to rgba8 (:726), DESIGN.md U5
  if (c != c) return 0;
  c = fminf(fmaxf(c, 0.0f), 1.0f);
  return (unsigned char)floorf(fadd(fmul(c, 255.0f), 0.5f));
}

// progressive running mean (:712-719, commented out upstream; svo_frame.flags bit 0)
SVO_DI float accumulate1(int frameNumber, unsigned char last8, float c) {
  const float last = fdiv((float)last8, 255.0f);  // imageLoad of the rgba8 image
  if (frameNumber < 100) return fdiv(fadd(fmul((float)frameNumber, last), c), (float)(frameNumber + 1));  // MAX_FRAME_ITER :43
  return last;
}
SVO_DI void pixel_finish(const SceneView &sc, const FrameParams &f, const Planes &pl, size_t p, Pixel &P) {
  if (P.x < 10 && P.y < 10)  // :696-700
    P.color = sc.first_word_zero ? mk3(1.0f, 0.0f, 0.0f) : mk3(1.0f, 1.0f, 1.0f);
  if ((f.flags & 1) && f.frameNumber > 1) {
    const uchar4 last = pl.rgba8[p];
    P.color = mk3(accumulate1(f.frameNumber, last.x, P.color.x), accumulate1(f.frameNumber, last.y, P.color.y),
                  accumulate1(f.frameNumber, last.z, P.color.z));
  }
}

// the end of main() (:696-727)
template <bool AUX>
SVO_DI void pixel_store(const SceneView &sc, const FrameParams &f, const Planes &pl, int W, Pixel &P) {
  const size_t p = (size_t)P.y * (size_t)W + (size_t)P.x;
  pixel_finish(sc, f, pl, p, P);
  pl.rgba8[p] = make_uchar4(quant8(P.color.x), quant8(P.color.y), quant8(P.color.z), 255);  // :726
  pl.depth[p] = P.depth;                                                                       // :727
  if (AUX) {
    pl.hit_id[p] = P.hit_id;
    pl.iter[p] = P.iter;
    pl.primary_t[p] = P.primary_t;
    pl.radiance[p] = make_float4(P.color.x, P.color.y, P.color.z, 1.0f);
  }
}

// main (svotrace.comp:649-729) for pixel (x, y), run to completion.  STACK: 0 default, 1 WideStack, 2 SmemStack (`smem_stack` =
// the thread's column of the CTA's shared-memory stack).
template <bool FAST, bool AUX, int STATS = 0, bool BOX = false, bool TOP = false, int STACK = 0, int BAL = 0>
__device__ __forceinline__ void shade_pixel(const SceneView &sc, const FrameParams &f, const Planes &pl, int W, int H,
                                            int x, int y, RayStats *rs = nullptr, uint2 *smem_stack = nullptr) {
  Pixel P;
  if (pixel_begin(f, pl, W, H, x, y, P)) {
    bool more;
    do {
      uint32_t loops = 0;
      const bool attrs = cast_needs_attrs(f, P);
      bool hit;
      // the primary cast may start at the beam pre-pass's lower bound; a block the pre-pass proved empty is not cast at all
      const float t_floor = (BOX && P.cast_i == 0) ? P.t_floor : 0.0f;
      if (BOX && t_floor == __uint_as_float(0x7f800000u)) {
        hit = false;
      } else
      if (STACK == 1) {
        uint4 wide[kMaxScale + 1];
        WideStack ws;
        ws.p = wide;
        hit = cast_ray_on<FAST, STATS, BOX, TOP, BAL>(ws, sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, P.res, loops, rs, attrs, t_floor);
      } else if (STACK == 2) {
        SmemStack ss;
        ss.p = smem_stack;
        hit = cast_ray_on<FAST, STATS, BOX, TOP, BAL>(ss, sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, P.res, loops, rs, attrs, t_floor);
      } else {
        hit = cast_ray<FAST, STATS, BOX, TOP>(sc, P.origin, P.dir, f.maxDepth, P.cone, f.coneDepth, P.res, loops, rs, attrs, t_floor);
      }
      if (!attrs && hit && f.renderMode == 0) {
        pixel_after_last_hit(P);
        more = false;
      } else {
        more = pixel_after_cast(f, P, hit, loops);
      }
    } while (more);
  }
  pixel_store<AUX>(sc, f, pl, W, P);
}

}  // namespace svo
