// rt_math.h — scalar/vector arithmetic with EXPLICIT rounding and fusion, host + device.
//
// Why explicit: parity with the reference is defined bit-for-bit (reference-RNG mode, primary-hit
// IDs). The reference is built with nvcc's default -fmad=true, so which multiplies fuse with which
// adds is decided by NVVM and ptxas, not by the source. We read those decisions off the SASS of the
// reference's sm_100 build (nvcc 12.9, `nvdisasm -gi`) and write them down here and at every call
// site with the non-contractable intrinsics (__fmaf_rn / __fmul_rn / __fadd_rn ...), so that this
// code computes the same bits no matter how it is inlined or scheduled. On the host the same
// functions use fmaf() and plain operators; host translation units are compiled with
// -ffp-contract=off.
//
// Contraction rules observed in the reference SASS (vec3.cuh operators after inlining):
//   a*b + c*d        -> fma(a, b, c*d)            (left product fused, right product rounded)
//   x + a*b, a*b + x -> fma(a, b, x)
//   x - a*b          -> fma(-a, b, x)
//   a*b - c*d        -> fma(a, b, -(c*d))
//   -a*b + c*d       -> fma(c, d, -(a*b))
//   dot(a,b) = a0*b0 + a1*b1 + a2*b2 -> fma(a2, b2, fma(a0, b0, a1*b1))        (vec3.cuh:92-95)
//   cross(a,b).x = fma(a1, b2, -(a2*b1)); .y = -fma(a0, b2, -(a2*b0)); .z = fma(a0, b1, -(a1*b0))
//   A + t*B          -> fma(t, B, A)                                            (ray.cuh:16)
#pragma once
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__
#else
#define RT_HD inline
#define RT_D inline
#endif

namespace rt {

#if defined(__CUDA_ARCH__)
RT_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
RT_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
RT_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
RT_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
RT_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
RT_HD float frcp(float a) { return __frcp_rn(a); }
RT_HD float fsqrt(float a) { return __fsqrt_rn(a); }
RT_HD uint32_t f2u(float f) { return __float_as_uint(f); }
RT_HD float u2f(uint32_t u) { return __uint_as_float(u); }
#else
RT_HD float fmul(float a, float b) { return a * b; }
RT_HD float fadd(float a, float b) { return a + b; }
RT_HD float fsub(float a, float b) { return a - b; }
RT_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
RT_HD float fdiv(float a, float b) { return a / b; }
RT_HD float frcp(float a) { return 1.0f / a; }
RT_HD float fsqrt(float a) { return sqrtf(a); }
RT_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RT_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#endif

struct V3 {
  float x, y, z;
};

RT_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD V3 vadd(V3 a, V3 b) { return v3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }   // vec3.cuh:57
RT_HD V3 vsub(V3 a, V3 b) { return v3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }   // vec3.cuh:62
RT_HD V3 vmul(V3 a, V3 b) { return v3(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }   // vec3.cuh:67
RT_HD V3 vscale(float t, V3 a) { return v3(fmul(t, a.x), fmul(t, a.y), fmul(t, a.z)); }    // vec3.cuh:77
RT_HD V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
RT_HD V3 vdivs(V3 a, float t) { return v3(fdiv(a.x, t), fdiv(a.y, t), fdiv(a.z, t)); }      // vec3.cuh:82
// A + t*B with the product fused (ray.cuh:16 point_at_parameter and every "x + t*v" form)
RT_HD V3 vmad(float t, V3 b, V3 a) { return v3(ffma(t, b.x, a.x), ffma(t, b.y, a.y), ffma(t, b.z, a.z)); }
// a - t*b -> fma(-t, b, a)
RT_HD V3 vnmad(float t, V3 b, V3 a) { return v3(ffma(-t, b.x, a.x), ffma(-t, b.y, a.y), ffma(-t, b.z, a.z)); }
RT_HD float vdot(V3 a, V3 b) { return ffma(a.z, b.z, ffma(a.x, b.x, fmul(a.y, b.y))); }    // vec3.cuh:92
RT_HD float vsqlen(V3 a) { return vdot(a, a); }                                            // vec3.cuh:33
RT_HD float vlen(V3 a) { return fsqrt(vsqlen(a)); }                                        // vec3.cuh:32
RT_HD V3 vcross(V3 a, V3 b) {                                                              // vec3.cuh:97
  return v3(ffma(a.y, b.z, -fmul(a.z, b.y)), -ffma(a.x, b.z, -fmul(a.z, b.x)), ffma(a.x, b.y, -fmul(a.y, b.x)));
}
RT_HD V3 vunit(V3 a) { return vdivs(a, vlen(a)); }                                         // vec3.cuh:155
RT_HD float vget(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

}  // namespace rt
