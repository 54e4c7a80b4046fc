/*
 * glsl_shim.h -- TEST INFRASTRUCTURE.  The GLSL 4.30 vocabulary that
 * /root/reference/src/shaders/svotrace.comp and svobeam.comp use, as C++17, so
 * that g++ can compile the reference's own shader TEXT (oracle/build_ref.py
 * reads it from /root/reference at build time, applies a mechanical rewrite and
 * writes the result under oracle/_ref/, which is git-ignored: no reference
 * source enters this repository).  The product never includes this file.
 *
 * What the shim decides is only what GLSL leaves to the GL driver -- the same
 * arithmetic contract as oracle_math.h (one IEEE binary32 rounding per + - * /,
 * minNum/maxNum, mix = x*(1-a)+y*a, dot = (xx'+yy')+zz', normalize = v/sqrt(dot),
 * oracle_math.h's sin/cos/acos/exp kernels) and the undefined-behaviour rules
 * U1 (out-of-range buffer/image reads return 0), U2 (uninitialised locals are
 * zero) and U5 (rgba8 store: NaN -> 0, clamp, floor(c*255+0.5)).  Control flow,
 * expression order, constants, record decode, traversal and shading come from
 * the shader text as the C++ parser reads it (GLSL and C++ agree on operator
 * precedence and associativity for every operator the shaders use).
 *
 * Compile with -ffp-contract=off -fno-fast-math.
 */
#ifndef SVO_GLSL_SHIM_H
#define SVO_GLSL_SHIM_H

#include <cstddef>
#include <cstdint>
#include <cstring>

#include "oracle_math.h"

namespace glsl {

typedef uint32_t uint;

/* ---- vectors ------------------------------------------------------------- */
struct vec2;
struct swz_xy { /* the `.xy` view of a vec2 (rand(): co.xy) */
  float a, b;
  operator vec2() const;
};
struct vec2 {
  union {
    struct { float x, y; };
    swz_xy xy;
  };
  vec2() : x(0.0f), y(0.0f) {}
  explicit vec2(float s) : x(s), y(s) {}
  vec2(float x_, float y_) : x(x_), y(y_) {}
  explicit vec2(const struct ivec2 &v);
};
inline swz_xy::operator vec2() const { return vec2(a, b); }

struct vec3 {
  union {
    struct { float x, y, z; };
    struct { float r, g, b; };
  };
  vec3() : x(0.0f), y(0.0f), z(0.0f) {}
  explicit vec3(float s) : x(s), y(s), z(s) {}
  vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

struct vec4 {
  union {
    struct { float x, y, z, w; };
    struct { float r, g, b, a; };
  };
  vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
  vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
  vec4(const vec3 &v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
};

struct uvec2 {
  uint x, y;
  uvec2() : x(0), y(0) {}
  uvec2(uint x_, uint y_) : x(x_), y(y_) {}
};
inline uvec2 operator*(const uvec2 &a, uint s) { return uvec2(a.x * s, a.y * s); }

struct ivec2 {
  int x, y;
  ivec2() : x(0), y(0) {}
  ivec2(int x_, int y_) : x(x_), y(y_) {}
  explicit ivec2(const uvec2 &v) : x((int)v.x), y((int)v.y) {}
};
inline ivec2 operator/(const ivec2 &a, int s) { return ivec2(a.x / s, a.y / s); }
inline vec2::vec2(const ivec2 &v) : x((float)v.x), y((float)v.y) {}

struct bvec2 { bool x, y; };
inline bvec2 greaterThanEqual(const ivec2 &a, const ivec2 &b) { return bvec2{a.x >= b.x, a.y >= b.y}; }
inline bool any(const bvec2 &v) { return v.x || v.y; }

struct uvec3_id { /* gl_GlobalInvocationID: only `.xy` is used */
  uvec2 xy;
};

/* component-wise arithmetic: one rounding per operation (-ffp-contract=off) */
inline vec2 operator+(const vec2 &a, const vec2 &b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator/(const vec2 &a, const vec2 &b) { return vec2(a.x / b.x, a.y / b.y); }

inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3 &a, const vec3 &b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator+(const vec3 &a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3 &a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(const vec3 &a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(float s, const vec3 &a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 &operator+=(vec3 &a, const vec3 &b) { a = a + b; return a; }
inline vec3 &operator-=(vec3 &a, const vec3 &b) { a = a - b; return a; }
inline vec3 &operator*=(vec3 &a, const vec3 &b) { a = a * b; return a; }
inline vec3 &operator+=(vec3 &a, float s) { a = a + s; return a; }
inline vec3 &operator*=(vec3 &a, float s) { a = a * s; return a; }

/* ---- built-in functions (oracle_math.h's contract) ------------------------ */
inline float abs(float a) { return om_abs(a); }
inline float sign(float a) { return om_sign(a); }
inline float min(float a, float b) { return om_min(a, b); }
inline float max(float a, float b) { return om_max(a, b); }
inline float fract(float a) { return om_fract(a); }
inline float sin(float a) { return om_sin(a); }
inline float cos(float a) { return om_cos(a); }
inline float acos(float a) { return om_acos(a); }
inline float exp(float a) { return om_exp(a); }
inline float sqrt(float a) { return sqrtf(a); }
inline uint floatBitsToUint(float f) { return om_f2u(f); }
inline float uintBitsToFloat(uint u) { return om_u2f(u); }
inline int findMSB(uint v) { return v ? 31 - __builtin_clz(v) : -1; }

inline float dot(const vec2 &a, const vec2 &b) {
  float p0 = a.x * b.x, p1 = a.y * b.y;
  return p0 + p1;
}
inline float dot(const vec3 &a, const vec3 &b) {
  float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.z * b.z;
  float s = p0 + p1;
  return s + p2;
}
inline vec3 cross(const vec3 &a, const vec3 &b) {
  float t0 = a.y * b.z, t1 = a.z * b.y, t2 = a.z * b.x, t3 = a.x * b.z, t4 = a.x * b.y, t5 = a.y * b.x;
  return vec3(t0 - t1, t2 - t3, t4 - t5);
}
inline vec3 normalize(const vec3 &v) {
  float len = sqrtf(dot(v, v));
  return vec3(v.x / len, v.y / len, v.z / len);
}
inline float mix(float x, float y, float a) { return om_mix(x, y, a); }
inline vec3 mix(const vec3 &x, const vec3 &y, float a) {
  return vec3(om_mix(x.x, y.x, a), om_mix(x.y, y.y, a), om_mix(x.z, y.z, a));
}

/* ---- buffers and images ---------------------------------------------------- */
struct ssbo_uint { /* `uint[] octreeBuffer`: little-endian word view of the byte stream, U1 */
  const uint8_t *bytes = nullptr;
  uint64_t nbytes = 0;
  uint operator[](uint i) const {
    uint64_t p = (uint64_t)i * 4u;
    if (p + 4u <= nbytes) {
      uint w;
      std::memcpy(&w, bytes + p, 4);
      return w;
    }
    uint w = 0;
    for (uint64_t k = 0; k < 4u && p + k < nbytes; k++) w |= (uint)bytes[p + k] << (8u * (unsigned)k);
    return w;
  }
};
struct ssbo_uint_rw { /* svobeam.comp's requestBuffer (only requestNode(), never called, writes it) */
  uint scratch[16] = {};
  uint &operator[](uint i) { return scratch[i & 15u]; }
};
struct isamplerBuffer {};

enum image_format { IMG_RGBA8, IMG_R32F, IMG_R32UI };
struct image2D {
  int width = 0, height = 0;
  image_format format = IMG_R32F;
  uint8_t *rgba8 = nullptr; /* IMG_RGBA8: 4 bytes per texel */
  float *f32 = nullptr;     /* IMG_R32F: 1 float per texel;  IMG_RGBA8: optional pre-quantisation copy, 4 floats */
};
typedef image2D uimage2D;

inline ivec2 imageSize(const image2D &img) { return ivec2(img.width, img.height); }
inline uint8_t quant_unorm8(float c) { /* U5 */
  if (c != c) return 0;
  if (c < 0.0f) c = 0.0f;
  if (c > 1.0f) c = 1.0f;
  return (uint8_t)floorf(c * 255.0f + 0.5f);
}
inline vec4 imageLoad(const image2D &img, const ivec2 &p) {
  if (p.x < 0 || p.y < 0 || p.x >= img.width || p.y >= img.height) return vec4(); /* U1 */
  size_t i = (size_t)p.y * (size_t)img.width + (size_t)p.x;
  if (img.format == IMG_R32F) return img.f32 ? vec4(img.f32[i], 0.0f, 0.0f, 1.0f) : vec4();
  if (img.format == IMG_RGBA8 && img.rgba8)
    return vec4((float)img.rgba8[4 * i] / 255.0f, (float)img.rgba8[4 * i + 1] / 255.0f,
                (float)img.rgba8[4 * i + 2] / 255.0f, (float)img.rgba8[4 * i + 3] / 255.0f);
  return vec4();
}
inline void imageStore(image2D &img, const ivec2 &p, const vec4 &v) {
  if (p.x < 0 || p.y < 0 || p.x >= img.width || p.y >= img.height) return;
  size_t i = (size_t)p.y * (size_t)img.width + (size_t)p.x;
  if (img.format == IMG_R32F) {
    if (img.f32) img.f32[i] = v.x;
    return;
  }
  if (img.format == IMG_RGBA8) {
    if (img.f32) { img.f32[4 * i] = v.x; img.f32[4 * i + 1] = v.y; img.f32[4 * i + 2] = v.z; img.f32[4 * i + 3] = v.w; }
    if (img.rgba8) {
      img.rgba8[4 * i] = quant_unorm8(v.x);
      img.rgba8[4 * i + 1] = quant_unorm8(v.y);
      img.rgba8[4 * i + 2] = quant_unorm8(v.z);
      img.rgba8[4 * i + 3] = quant_unorm8(v.w);
    }
  }
}

}  // namespace glsl
#endif
