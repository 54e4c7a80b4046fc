// detmath.cuh -- the fp32 arithmetic contract of the trace path on the device.
//
// GLSL leaves sin/cos/acos/exp precision, fma contraction and the expansion of
// mix/normalize to the GL driver (reference: src/shaders/svotrace.comp uses
// sin :27, cos/sin :506, acos :547, exp :596-598, normalize :387,496,506,675,
// mix :664).  This build fixes one interpretation and implements it with
// explicitly rounded intrinsics, so the result does not depend on --fmad or
// --use_fast_math:
//   every add/sub/mul is its own IEEE binary32 rounding, div and sqrt are
//   correctly rounded, min/max ignore a NaN operand (FMNMX, what NVIDIA's GLSL
//   compiler emits), mix(x,y,a) = x*(1-a) + y*a, dot = (x*x' + y*y') + z*z',
//   normalize(v) = v / sqrt(dot(v,v)).
// Transcendentals are fixed polynomial kernels (Cephes single-precision
// coefficients) with a binary64 Cody-Waite reduction for sin/cos: rand()
// (svotrace.comp:26-29) feeds sin() arguments near 1e5..1e6, where the
// reduction decides every bit of the result.
#pragma once
#include <cstdint>

namespace svo {

#define SVO_DI __device__ __forceinline__

SVO_DI float fmul(float a, float b) { return __fmul_rn(a, b); }
SVO_DI float fadd(float a, float b) { return __fadd_rn(a, b); }
SVO_DI float fsub(float a, float b) { return __fsub_rn(a, b); }
SVO_DI float fdiv(float a, float b) { return __fdiv_rn(a, b); }
SVO_DI float fsqrt(float a) { return __fsqrt_rn(a); }

struct vec3 {
  float x, y, z;
};
SVO_DI vec3 mk3(float x, float y, float z) { vec3 v; v.x = x; v.y = y; v.z = z; return v; }

SVO_DI float mixf(float x, float y, float a) { return fadd(fmul(x, fsub(1.0f, a)), fmul(y, a)); }
SVO_DI float dot3(vec3 a, vec3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
SVO_DI vec3 normalize3(vec3 v) {
  float len = fsqrt(dot3(v, v));
  return mk3(fdiv(v.x, len), fdiv(v.y, len), fdiv(v.z, len));
}
SVO_DI vec3 cross3(vec3 a, vec3 b) {
  return mk3(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
             fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
SVO_DI float qnan() { return __uint_as_float(0x7fc00000u); }

// r in [-pi/4, pi/4], quadrant in q.  pi/2 = P1 + P2 with P1 carrying 33 bits,
// so k*P1 is exact in binary64 for |k| < 2^20.
SVO_DI float reduce_pio2(float x, int &q) {
  double xd = (double)x;
  double kd = rint(__dmul_rn(xd, 0.63661977236758138243));
  double r = __dsub_rn(xd, __dmul_rn(kd, 1.57079632673412561417e+00));
  r = __dsub_rn(r, __dmul_rn(kd, 6.07710050650619224932e-11));
  q = (int)((long long)kd & 3);
  return __double2float_rn(r);
}
SVO_DI float sin_poly(float r) {
  float z = fmul(r, r);
  float p = fmul(-1.9515295891e-4f, z);
  p = fadd(p, 8.3321608736e-3f);
  p = fmul(p, z);
  p = fadd(p, -1.6666654611e-1f);
  p = fmul(p, z);
  p = fmul(p, r);
  return fadd(p, r);
}
SVO_DI float cos_poly(float r) {
  float z = fmul(r, r);
  float p = fmul(2.443315711809948e-5f, z);
  p = fadd(p, -1.388731625493765e-3f);
  p = fmul(p, z);
  p = fadd(p, 4.166664568298827e-2f);
  p = fmul(p, z);
  p = fmul(p, z);
  p = fsub(p, fmul(0.5f, z));
  return fadd(p, 1.0f);
}
SVO_DI float det_sin(float x) {
  if (!(fabsf(x) < 1.0e9f)) return qnan();
  int q;
  float r = reduce_pio2(x, q);
  float s = (q & 1) ? cos_poly(r) : sin_poly(r);
  return (q & 2) ? -s : s;
}
SVO_DI float det_cos(float x) {
  if (!(fabsf(x) < 1.0e9f)) return qnan();
  int q;
  float r = reduce_pio2(x, q);
  float c = (q & 1) ? sin_poly(r) : cos_poly(r);
  return ((q + 1) & 2) ? -c : c;
}

SVO_DI float asin_core(float a) {
  float z = fmul(a, a);
  float p = fmul(4.2163199048e-2f, z);
  p = fadd(p, 2.4181311049e-2f);
  p = fmul(p, z);
  p = fadd(p, 4.5470025998e-2f);
  p = fmul(p, z);
  p = fadd(p, 7.4953002686e-2f);
  p = fmul(p, z);
  p = fadd(p, 1.6666752422e-1f);
  p = fmul(p, z);
  p = fmul(p, a);
  return fadd(p, a);
}
SVO_DI float det_acos(float x) {
  if (!(fabsf(x) <= 1.0f)) return qnan();
  if (x < -0.5f) {
    float s = asin_core(fsqrt(fmul(0.5f, fadd(1.0f, x))));
    return fsub(3.14159265358979323846f, fmul(2.0f, s));
  }
  if (x > 0.5f) {
    float s = asin_core(fsqrt(fmul(0.5f, fsub(1.0f, x))));
    return fmul(2.0f, s);
  }
  return fsub(1.57079632679489661923f, asin_core(x));
}

SVO_DI float det_exp(float x) {
  if (x != x) return x;
  if (x > 88.7f) return __uint_as_float(0x7f800000u);
  if (x < -103.0f) return 0.0f;
  float kf = rintf(fmul(x, 1.44269504088896341f));
  float r = fsub(x, fmul(kf, 0.693359375f));
  r = fsub(r, fmul(kf, -2.12194440e-4f));
  float z = fmul(r, r);
  float p = fmul(1.9875691500e-4f, r);
  p = fadd(p, 1.3981999507e-3f);
  p = fmul(p, r);
  p = fadd(p, 8.3334519073e-3f);
  p = fmul(p, r);
  p = fadd(p, 4.1665795894e-2f);
  p = fmul(p, r);
  p = fadd(p, 1.6666665459e-1f);
  p = fmul(p, r);
  p = fadd(p, 5.0000001201e-1f);
  p = fmul(p, z);
  p = fadd(p, r);
  p = fadd(p, 1.0f);
  int k = (int)kf;
  if (k < -126) {
    p = fmul(p, __uint_as_float((uint32_t)(k + 64 + 127) << 23));
    return fmul(p, __uint_as_float((uint32_t)(-64 + 127) << 23));
  }
  if (k > 127) {
    p = fmul(p, __uint_as_float((uint32_t)(k - 64 + 127) << 23));
    return fmul(p, __uint_as_float((uint32_t)(64 + 127) << 23));
  }
  return fmul(p, __uint_as_float((uint32_t)(k + 127) << 23));
}

// svotrace.comp:26-29
SVO_DI float det_rand(float x, float y) {
  float s = det_sin(fadd(fmul(x, 12.9898f), fmul(y, 78.233f)));
  float m = fmul(s, 43758.5453f);
  return fsub(m, floorf(m));
}

}  // namespace svo
