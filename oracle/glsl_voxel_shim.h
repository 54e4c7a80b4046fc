// glsl_voxel_shim.h -- TEST INFRASTRUCTURE.  The GLSL vocabulary of the reference's voxeliser,
// src/shaders/chunkgen-heightmap.comp (integer images, ivec2/3/4, gl_GlobalInvocationID swizzles), as C++, so that the
// shader's text -- rewritten mechanically by oracle/build_ref_java.py (rules R1-R3 of oracle/build_ref.py), never copied
// into the repository -- compiles for the CPU and feeds the compiled Java builder the voxels the engine's GL dispatch would.
// Images follow GL: imageLoad/imageStore outside the image read 0 / are dropped; r16ui loads are unsigned, r8i loads and
// stores signed 8-bit.
#pragma once
#include <cstddef>
#include <cstdint>

namespace glslv {

typedef uint32_t uint;
#define readonly
#define writeonly

struct ivec2 {
  int x = 0, y = 0;
  ivec2() {}
  ivec2(int x_, int y_) : x(x_), y(y_) {}
  template <class V>
  explicit ivec2(const V &v) : x((int)v.x), y((int)v.y) {}
};
inline ivec2 operator+(const ivec2 &a, const ivec2 &b) { return ivec2(a.x + b.x, a.y + b.y); }
struct ivec3 {
  int x = 0, y = 0, z = 0;
  ivec3() {}
  ivec3(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
  template <class V>
  explicit ivec3(const V &v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
};
struct ivec4 {
  int r = 0, g = 0, b = 0, a = 0;
  ivec4() {}
  template <class A>
  ivec4(A r_, int g_, int b_, int a_) : r((int)r_), g(g_), b(b_), a(a_) {}
};
struct uvec4 {
  uint r = 0, g = 0, b = 0, a = 0;
};

// gl_GlobalInvocationID with the swizzles the shader spells: .y, .xz, .xyz
struct global_invocation_id {
  uint x = 0, y = 0, z = 0;
  struct { uint x, y; } xz = {0, 0};
  struct { uint x, y, z; } xyz = {0, 0, 0};
  void set(uint x_, uint y_, uint z_) {
    x = x_; y = y_; z = z_;
    xz.x = x_; xz.y = z_;
    xyz.x = x_; xyz.y = y_; xyz.z = z_;
  }
};

struct uimage2D {  // r16ui
  const uint16_t *texels = nullptr;
  int width = 0, height = 0;
};
struct iimage2D {  // r8i
  const int8_t *texels = nullptr;
  int width = 0, height = 0;
};
struct iimage3D {  // r8i; pitch in texels between rows / slices (the engine reads it back into x | y << 10 | z << 20)
  int8_t *texels = nullptr;
  int width = 0, height = 0, depth = 0;
  size_t row_pitch = 0, slice_pitch = 0;
};
inline uvec4 imageLoad(const uimage2D &img, const ivec2 &p) {
  uvec4 v;
  if (p.x >= 0 && p.y >= 0 && p.x < img.width && p.y < img.height) v.r = img.texels[(size_t)p.y * (size_t)img.width + (size_t)p.x];
  return v;
}
inline ivec4 imageLoad(const iimage2D &img, const ivec2 &p) {
  ivec4 v;
  if (p.x >= 0 && p.y >= 0 && p.x < img.width && p.y < img.height) v.r = img.texels[(size_t)p.y * (size_t)img.width + (size_t)p.x];
  return v;
}
inline void imageStore(iimage3D &img, const ivec3 &p, const ivec4 &v) {
  if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= img.width || p.y >= img.height || p.z >= img.depth) return;
  const int c = v.r < -128 ? -128 : v.r > 127 ? 127 : v.r;  // r8i stores clamp to the format's range
  img.texels[(size_t)p.x + (size_t)p.y * img.row_pitch + (size_t)p.z * img.slice_pitch] = (int8_t)c;
}

struct invocation_base {
  global_invocation_id gl_GlobalInvocationID;
};

}  // namespace glslv
