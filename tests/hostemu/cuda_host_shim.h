// cuda_host_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile the DEVICE source of the trace path
// (svo_raytracer_b200/csrc/svo_trace.cuh + detmath.cuh) for the host, so that the exact statements the sm_100a
// kernels execute can be checked against the oracle in the CPU test suite (no GPU in the build container) and
// driven warp by warp by the SIMT divergence model (tests/hostemu/emu.cpp).  Nothing here is linked into
// libsvo_b200.so; the product has no CPU path.
//
// Every shim is the documented semantics of the intrinsic it stands for (CUDA Math API / PTX ISA):
//   __f{add,sub,mul,div}_rn, __fsqrt_rn : one IEEE-754 binary32 operation, round to nearest even, never
//        contracted (built with -ffp-contract=off; x86-64 SSE arithmetic is IEEE);
//   fminf / fmaxf : return the non-NaN operand (device: FMNMX; host: C99 fmin/fmax) -- same rule;
//   __popc, __funnelshift_r, bfind.u32 (find_msb's #else branch), __float_as_uint ... : bit operations.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>  // uint2, uchar4, float4, make_* (host definitions); __device__ expands to nothing under g++

#ifdef __CUDACC__
#error "cuda_host_shim.h is for plain g++ builds of the device headers"
#endif

static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
  return (uint32_t)((((uint64_t)hi << 32) | (uint64_t)lo) >> (shift & 31u));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t shift) {
  return (uint32_t)(((((uint64_t)hi << 32) | (uint64_t)lo) << (shift & 31u)) >> 32);
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
