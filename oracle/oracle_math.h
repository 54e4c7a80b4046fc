/*
 * oracle_math.h -- TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
 *
 * Deterministic fp32 arithmetic contract for the restatement of
 * /root/reference/src/shaders/svotrace.comp.
 *
 * GLSL leaves fma contraction, mix/normalize expansion and the precision of
 * sin/cos/acos/exp to the GL driver.  The oracle therefore FIXES one
 * interpretation (stated below); the CUDA validation build implements the
 * same interpretation independently (svo_raytracer_b200/csrc/detmath.cuh)
 * and the two are compared bit-for-bit in tests/.
 *
 *   - every + - * is a separately rounded IEEE-754 binary32 operation
 *     (compile with -ffp-contract=off, no -ffast-math);
 *   - division and sqrt are correctly rounded;
 *   - min/max are IEEE minNum/maxNum (fminf/fmaxf: a NaN operand is ignored)
 *     -- this is what NVIDIA's GLSL compiler emits (FMNMX) and it is what
 *     makes the reference's axis-parallel rays (t_coef = -inf, NaN corners)
 *     behave as they do on the author's hardware;
 *   - mix(x,y,a) = x*(1-a) + y*a            (GLSL 4.30 spec, section 8.3);
 *   - dot(a,b)  = (a.x*b.x + a.y*b.y) + a.z*b.z;
 *   - normalize(v) = v / sqrt(dot(v,v))     (component-wise IEEE division);
 *   - sin/cos/acos/exp: the fixed polynomial kernels below (Cephes single
 *     precision coefficients, public domain), range reduction for sin/cos in
 *     binary64 so that rand()'s arguments (up to ~1e6) reduce exactly.
 */
#ifndef SVO_ORACLE_MATH_H
#define SVO_ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t om_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float om_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline float om_min(float a, float b) { return fminf(a, b); }
static inline float om_max(float a, float b) { return fmaxf(a, b); }
static inline float om_abs(float a) { return fabsf(a); }

/* GLSL sign(): 1, -1, or 0 (also 0 for -0; NaN never reaches it, see caller) */
static inline float om_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

static inline float om_mix(float x, float y, float a) {
  float ia = 1.0f - a;
  float xa = x * ia;
  float ya = y * a;
  return xa + ya;
}

static inline float om_dot3(const float a[3], const float b[3]) {
  float p0 = a[0] * b[0];
  float p1 = a[1] * b[1];
  float p2 = a[2] * b[2];
  float s = p0 + p1;
  return s + p2;
}

static inline void om_normalize3(const float v[3], float out[3]) {
  float len = sqrtf(om_dot3(v, v));
  out[0] = v[0] / len;
  out[1] = v[1] / len;
  out[2] = v[2] / len;
}

/* cross(a,b) = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x) */
static inline void om_cross3(const float a[3], const float b[3], float out[3]) {
  float t0 = a[1] * b[2], t1 = a[2] * b[1];
  float t2 = a[2] * b[0], t3 = a[0] * b[2];
  float t4 = a[0] * b[1], t5 = a[1] * b[0];
  out[0] = t0 - t1;
  out[1] = t2 - t3;
  out[2] = t4 - t5;
}

static inline float om_fract(float x) { return x - floorf(x); }

/* ---- sin / cos ---------------------------------------------------------- */
/* reduce x to r in [-pi/4, pi/4] and quadrant q; binary64 Cody-Waite with the
 * fdlibm split of pi/2 (pio2_1 has 33 significant bits: k*pio2_1 is exact for
 * |k| < 2^20). */
static inline float om_reduce_pio2(float x, int *q) {
  double xd = (double)x;
  double kd = nearbyint(xd * 0.63661977236758138243); /* round-half-even */
  double p1 = kd * 1.57079632673412561417e+00;
  double p2 = kd * 6.07710050650619224932e-11;
  double r = xd - p1;
  r = r - p2;
  *q = (int)((long long)kd & 3);
  return (float)r;
}
static inline float om_sin_poly(float r) {
  float z = r * r;
  float p = -1.9515295891e-4f * z;
  p = p + 8.3321608736e-3f;
  p = p * z;
  p = p + -1.6666654611e-1f;
  p = p * z;
  p = p * r;
  return p + r;
}
static inline float om_cos_poly(float r) {
  float z = r * r;
  float p = 2.443315711809948e-5f * z;
  p = p + -1.388731625493765e-3f;
  p = p * z;
  p = p + 4.166664568298827e-2f;
  p = p * z;
  p = p * z;
  float h = 0.5f * z;
  p = p - h;
  return p + 1.0f;
}
static inline float om_sin(float x) {
  if (!(fabsf(x) < 1.0e9f)) return om_u2f(0x7fc00000u);
  int q;
  float r = om_reduce_pio2(x, &q);
  float s = (q & 1) ? om_cos_poly(r) : om_sin_poly(r);
  return (q & 2) ? -s : s;
}
static inline float om_cos(float x) {
  if (!(fabsf(x) < 1.0e9f)) return om_u2f(0x7fc00000u);
  int q;
  float r = om_reduce_pio2(x, &q);
  float c = (q & 1) ? om_sin_poly(r) : om_cos_poly(r);
  return ((q + 1) & 2) ? -c : c;
}

/* ---- acos ---------------------------------------------------------------- */
static inline float om_asin_core(float a) { /* |a| <= 0.5 */
  float z = a * a;
  float p = 4.2163199048e-2f * z;
  p = p + 2.4181311049e-2f;
  p = p * z;
  p = p + 4.5470025998e-2f;
  p = p * z;
  p = p + 7.4953002686e-2f;
  p = p * z;
  p = p + 1.6666752422e-1f;
  p = p * z;
  p = p * a;
  return p + a;
}
static inline float om_acos(float x) {
  if (!(fabsf(x) <= 1.0f)) return om_u2f(0x7fc00000u);
  if (x < -0.5f) {
    float h = 1.0f + x;
    h = 0.5f * h;
    float s = om_asin_core(sqrtf(h));
    s = 2.0f * s;
    return 3.14159265358979323846f - s;
  }
  if (x > 0.5f) {
    float h = 1.0f - x;
    h = 0.5f * h;
    float s = om_asin_core(sqrtf(h));
    return 2.0f * s;
  }
  return 1.57079632679489661923f - om_asin_core(x);
}

/* ---- exp ----------------------------------------------------------------- */
static inline float om_exp(float x) {
  if (x != x) return x;
  if (x > 88.7f) return om_u2f(0x7f800000u);
  if (x < -103.0f) return 0.0f;
  float kf = nearbyintf(x * 1.44269504088896341f);
  float r = x - kf * 0.693359375f;
  r = r - kf * -2.12194440e-4f;
  float z = r * r;
  float p = 1.9875691500e-4f * r;
  p = p + 1.3981999507e-3f;
  p = p * r;
  p = p + 8.3334519073e-3f;
  p = p * r;
  p = p + 4.1665795894e-2f;
  p = p * r;
  p = p + 1.6666665459e-1f;
  p = p * r;
  p = p + 5.0000001201e-1f;
  p = p * z;
  p = p + r;
  p = p + 1.0f;
  int k = (int)kf;
  if (k < -126) { /* two-step scaling keeps the subnormal result correctly formed */
    p = p * om_u2f((uint32_t)(k + 64 + 127) << 23);
    return p * om_u2f((uint32_t)(-64 + 127) << 23);
  }
  if (k > 127) {
    p = p * om_u2f((uint32_t)(k - 64 + 127) << 23);
    return p * om_u2f((uint32_t)(64 + 127) << 23);
  }
  return p * om_u2f((uint32_t)(k + 127) << 23);
}

#endif
