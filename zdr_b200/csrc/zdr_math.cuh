// zdr_math.cuh — float3 algebra of the render kernels.
//
// NUMERIC CONTRACT: the library is compiled with -fmad=false, so every
// float operation below is a single IEEE-754 binary32 operation in source order; fused
// multiply-adds happen ONLY where fmaf() is written.  dot/cross/lincomb/lerp/mat-vec use the
// fma chains below; x**2 = x*x, x**5 = (x*x)*(x*x)*x; normalize(v) = v * (1/sqrt(v.v)).
// LuisaCompute's own lowering of these builtins is not inspectable, so this
// contract *defines* them for the framework; the CPU oracle states the same contract.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef ZDR_FAST_SINCOS
#define ZDR_FAST_SINCOS 0
#endif
#define ZDR_DEV __device__ __forceinline__
#define ZDR_HD __host__ __device__ __forceinline__

namespace zdr {

constexpr float kPi = 3.14159274101257324f;       // float(math.pi)
constexpr float kTwoPi = 6.28318548202514648f;    // float(2*math.pi)

ZDR_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
ZDR_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
ZDR_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
ZDR_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
ZDR_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
ZDR_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
ZDR_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
ZDR_HD float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

ZDR_HD float dot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
ZDR_HD float3 cross(float3 a, float3 b) {
    return f3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
// v * (1/sqrt(v.v)).  On the device the IEEE sqrt + divide expand to ~45 instructions with their slow
// paths; ~20 call sites per path vertex made this the largest single contributor to the render
// kernels' code size (ncu source page, profiles/r01g), so the device version is ONE out-of-line copy.
#if defined(__CUDA_ARCH__) && defined(ZDR_FAST_MATH)
// "fast" build of the shading code (Makefile: --use_fast_math): one MUFU.RSQ, inline
ZDR_HD float3 normalize(float3 v) { return v * rsqrtf(fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x))); }
#elif defined(__CUDA_ARCH__)
#ifndef ZDR_INLINE_RCP_LENGTH      // 1: the IEEE sqrt + divide inline at every call site (no call / register saves, more code).  Set per
#define ZDR_INLINE_RCP_LENGTH 0      // translation unit: the wavefront / persistent stages gain 0.8 / 3 % inline, the per-pixel kernels lose 2-3 % (r03b)
#endif
#if ZDR_INLINE_RCP_LENGTH
static __device__ __forceinline__
#else
static __device__ __noinline__
#endif
float rcp_length(float x, float y, float z) { return 1.0f / sqrtf(fmaf(z, z, fmaf(y, y, x * x))); }
ZDR_HD float3 normalize(float3 v) { return v * rcp_length(v.x, v.y, v.z); }
#else
ZDR_HD float3 normalize(float3 v) {
    float inv = 1.0f / sqrtf(dot(v, v));
    return v * inv;
}
#endif
ZDR_HD float length(float3 v) { return sqrtf(dot(v, v)); }
// a*A + b*B + c*C
ZDR_HD float3 lincomb(float a, float3 A, float b, float3 B, float c, float3 C) {
    return f3(fmaf(c, C.x, fmaf(b, B.x, a * A.x)), fmaf(c, C.y, fmaf(b, B.y, a * A.y)),
              fmaf(c, C.z, fmaf(b, B.z, a * A.z)));
}
ZDR_HD float lerp(float a, float b, float t) { return fmaf(t, b - a, a); }
ZDR_HD float4 lerp(float4 a, float4 b, float t) {
    return make_float4(lerp(a.x, b.x, t), lerp(a.y, b.y, t), lerp(a.z, b.z, t), lerp(a.w, b.w, t));
}
ZDR_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
ZDR_HD int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
ZDR_HD bool any_nan(float3 a) { return (a.x != a.x) || (a.y != a.y) || (a.z != a.z); }
ZDR_HD bool any_nan(float4 a) { return (a.x != a.x) || (a.y != a.y) || (a.z != a.z) || (a.w != a.w); }
// May this texel gradient be scattered?  The reference drops NaN (prb.py:178-187); the "fast" build also drops +-inf, which
// approximate divisions by nearly-zero BRDF values produce where the IEEE ones give huge finite numbers.
ZDR_HD bool grad_ok(float4 g) {
#ifdef ZDR_FAST_MATH
    return fabsf(g.x) <= 3.0e38f && fabsf(g.y) <= 3.0e38f && fabsf(g.z) <= 3.0e38f && fabsf(g.w) <= 3.0e38f;
#else
    return !any_nan(g);
#endif
}
ZDR_HD bool any_pos(float3 a) { return a.x > 0.0f || a.y > 0.0f || a.z > 0.0f; }

// float -> int, saturating, NaN -> 0 (cvt.rzi.s32.f32)
ZDR_DEV int f2i(float x) { return __float2int_rz(x); }

// (M * float4(v,1)).xyz, M column-major
ZDR_HD float3 xform_point(const float* M, float3 v) {
    return f3(fmaf(M[8], v.z, fmaf(M[4], v.y, fmaf(M[0], v.x, M[12]))),
              fmaf(M[9], v.z, fmaf(M[5], v.y, fmaf(M[1], v.x, M[13]))),
              fmaf(M[10], v.z, fmaf(M[6], v.y, fmaf(M[2], v.x, M[14]))));
}
// 3x3 column-major times vector
ZDR_HD float3 mat3_mul(const float* N, float3 v) {
    return f3(fmaf(N[6], v.z, fmaf(N[3], v.y, N[0] * v.x)), fmaf(N[7], v.z, fmaf(N[4], v.y, N[1] * v.x)),
              fmaf(N[8], v.z, fmaf(N[5], v.y, N[2] * v.x)));
}

// sin/cos of a binary32 angle evaluated in binary64: Cody-Waite reduction by pi/2 and Taylor
// polynomials in fma-Horner form, rounded once.  Identical operation sequence on host and
// device, so the sampled directions do not depend on a vendor libm.
#ifndef ZDR_INLINE_SINCOS
#define ZDR_INLINE_SINCOS 0
#endif
#if defined(__CUDA_ARCH__) && ZDR_INLINE_SINCOS
static __device__ __forceinline__
#elif defined(__CUDA_ARCH__)
static __device__ __noinline__
#else
inline
#endif
void sincos_contract(float phi, float* s_out, float* c_out) {
#if defined(__CUDA_ARCH__) && defined(ZDR_FAST_MATH) && ZDR_FAST_SINCOS
    // MUFU.SIN / MUFU.COS (absolute error ~2^-21 on [0, 2 pi]).  Measured (profiles/r02u): the 5e-7 rad this puts on every
    // sampled direction moves the next hit by ~1e-4 units, which a white-noise texture turns into ~1e-4 relative changes of
    // the material — the path integrator's samples then differ from the oracle's by > 1e-3 in 10 % of the cases.  Off: the
    // fast build keeps the binary64 polynomial below (3 % of the shade stage).
    __sincosf(phi, s_out, c_out);
    return;
#endif
    double x = (double)phi;
    double kd = rint(x * 0.63661977236758134308);
    double r = fma(-kd, 1.57079632679489655800e+00, x);
    r = fma(-kd, 6.12323399573676603587e-17, r);
    double r2 = r * r;
    double ps = -1.0 / 1307674368000.0;
    ps = fma(ps, r2, 1.0 / 6227020800.0);
    ps = fma(ps, r2, -1.0 / 39916800.0);
    ps = fma(ps, r2, 1.0 / 362880.0);
    ps = fma(ps, r2, -1.0 / 5040.0);
    ps = fma(ps, r2, 1.0 / 120.0);
    ps = fma(ps, r2, -1.0 / 6.0);
    double sn = fma(ps * r2, r, r);
    double pc = 1.0 / 20922789888000.0;
    pc = fma(pc, r2, -1.0 / 87178291200.0);
    pc = fma(pc, r2, 1.0 / 479001600.0);
    pc = fma(pc, r2, -1.0 / 3628800.0);
    pc = fma(pc, r2, 1.0 / 40320.0);
    pc = fma(pc, r2, -1.0 / 720.0);
    pc = fma(pc, r2, 1.0 / 24.0);
    pc = fma(pc, r2, -0.5);
    double cs = fma(pc, r2, 1.0);
    int q = (int)((long long)kd & 3);
    double s = (q == 0) ? sn : (q == 1) ? cs : (q == 2) ? -sn : -cs;
    double c = (q == 0) ? cs : (q == 1) ? -sn : (q == 2) ? -cs : sn;
    *s_out = (float)s;
    *c_out = (float)c;
}

// Division of 0 <= n < 2^31 by a launch constant d >= 1 as one 32x32 -> 64-bit multiply and a shift:
// with s = ceil(log2 d) and m = ceil(2^(31+s) / d) < 2^32, floor(n * m / 2^(31+s)) = floor(n / d) exactly
// (m = 2^(31+s)/d + e, 0 <= e < 1, and n * e * d < 2^(31+s)).  A hardware-less integer division by a runtime
// value costs ~20 instructions; the wavefront stages do two per path per stage.
struct FastDiv {
    unsigned m;
    int shift, d;
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    int s = 0;
    while ((1ll << s) < (long long)d) ++s;
    f.shift = 31 + s;
    f.m = (unsigned)(((1ull << f.shift) + (unsigned long long)d - 1ull) / (unsigned long long)d);
    f.d = d;
    return f;
}
ZDR_HD int fastdiv(int n, const FastDiv& f) { return (int)(((unsigned long long)(unsigned)n * f.m) >> f.shift); }

}  // namespace zdr
