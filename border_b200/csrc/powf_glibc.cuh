// powf_glibc.cuh -- bit-exact restatement of glibc's powf for positive finite bases.
//
// Why: border's SumTree raises priorities with f32::powf (reference
// border-core/src/generic_replay_buffer/base/sum_tree.rs:76,96,134,139), which Rust lowers to the
// platform libm.  "Bit-exact priorities" therefore means bit-exact with glibc powf.  CUDA's powf
// is a different algorithm, so the device path restates glibc's: log2 via a 16-entry table and a
// degree-5 polynomial, exp2 via a 32-entry table and a degree-3 polynomial, all in double.
// Tables come from the host libm (tools/gen_powf_tables.py -> powf_tables.inc).
//
// glibc on x86-64 selects an FMA build of powf at run time (ifunc) when the CPU has FMA; in that
// build every a*b+c below is fused.  `fused` picks the variant (1 = FMA build, the default on any
// current x86-64 host; 0 = SSE2 build).
#pragma once
#include <stdint.h>
#include "powf_tables.inc"

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD static inline
#endif

namespace bbpow {

struct U64Pair { uint64_t invc, logc; };

#if defined(__CUDACC__)
__device__ __constant__ static U64Pair kLog2Tab[16] = {BB_POWF_LOG2_TAB};
__device__ __constant__ static uint64_t kLog2Poly[5] = {BB_POWF_LOG2_POLY};
__device__ __constant__ static uint64_t kExp2Tab[32] = {BB_EXP2F_TAB};
__device__ __constant__ static uint64_t kExp2Poly[3] = {BB_EXP2F_POLY};
#endif
static const U64Pair hLog2Tab[16] = {BB_POWF_LOG2_TAB};
static const uint64_t hLog2Poly[5] = {BB_POWF_LOG2_POLY};
static const uint64_t hExp2Tab[32] = {BB_EXP2F_TAB};
static const uint64_t hExp2Poly[3] = {BB_EXP2F_POLY};
#if defined(__CUDA_ARCH__)
#define BB_TAB_LOG2 kLog2Tab
#define BB_TAB_LOG2P kLog2Poly
#define BB_TAB_EXP2 kExp2Tab
#define BB_TAB_EXP2P kExp2Poly
#else
#define BB_TAB_LOG2 hLog2Tab
#define BB_TAB_LOG2P hLog2Poly
#define BB_TAB_EXP2 hExp2Tab
#define BB_TAB_EXP2P hExp2Poly
#endif

BB_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; __builtin_memcpy(&d, &u, 8); return d;
#endif
}
BB_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; __builtin_memcpy(&u, &d, 8); return u;
#endif
}
BB_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; __builtin_memcpy(&u, &f, 4); return u;
#endif
}
BB_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}
// Explicitly rounded primitives so neither nvcc nor gcc contracts or reassociates.
BB_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
BB_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
BB_HD double dfma(double a, double b, double c, int fused) {
#if defined(__CUDA_ARCH__)
    return fused ? __fma_rn(a, b, c) : __dadd_rn(__dmul_rn(a, b), c);
#else
    return fused ? __builtin_fma(a, b, c) : dadd(dmul(a, b), c);
#endif
}

// powf(x, y) for the arguments SumTree produces: x >= 0 finite, y finite.  Other inputs take the
// closest IEEE answer (they never occur on the replay path and make the reference panic).
BB_HD float powf_glibc(float x, float y, int fused) {
    uint32_t ix = f2u(x), iy = f2u(y);
    if ((iy << 1) == 0) return 1.0f;                  // y == +-0
    if (ix == 0x3f800000u) return 1.0f;               // x == 1
    if ((ix << 1) == 0) return (iy >> 31) ? u2f(0x7f800000u) : 0.0f;  // x == 0
    if (ix >= 0x7f800000u || (iy & 0x7fffffffu) >= 0x7f800000u) {
        // x negative/inf/nan or y inf/nan: not reachable from SumTree; return NaN-or-limit.
        if (ix == 0x7f800000u) return (iy >> 31) ? 0.0f : x;
        return u2f(0x7fc00000u);
    }
    if (ix < 0x00800000u) {  // subnormal x: normalise
        ix = f2u(x * 8388608.0f);
        ix &= 0x7fffffffu;
        ix -= 23u << 23;
    }
    // log2_inline
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> (23 - 4)) % 16);
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double invc = u2d(BB_TAB_LOG2[i].invc), logc = u2d(BB_TAB_LOG2[i].logc);
    double z = (double)u2f(iz);
    double A0 = u2d(BB_TAB_LOG2P[0]), A1 = u2d(BB_TAB_LOG2P[1]), A2 = u2d(BB_TAB_LOG2P[2]),
           A3 = u2d(BB_TAB_LOG2P[3]), A4 = u2d(BB_TAB_LOG2P[4]);
    double r = dfma(z, invc, -1.0, fused);
    double y0 = dadd(logc, (double)k);
    double r2 = dmul(r, r);
    double yy = dfma(A0, r, A1, fused);
    double p = dfma(A2, r, A3, fused);
    double r4 = dmul(r2, r2);
    double q = dfma(A4, r, y0, fused);
    q = dfma(p, r2, q, fused);
    yy = dfma(yy, r4, q, fused);
    double ylogx = dmul((double)y, yy);
    if (((d2u(ylogx) >> 47) & 0xffff) >= (d2u(126.0) >> 47)) {
        if (ylogx > u2d(0x405fffffffd1d571ULL)) return u2f(0x7f800000u);  // overflow
        if (ylogx <= -150.0) return 0.0f;                                  // underflow
    }
    // exp2_inline
    const double SHIFT = u2d(BB_EXP2F_SHIFT_SCALED);
    double kd = dadd(ylogx, SHIFT);
    uint64_t ki = d2u(kd);
    kd = dadd(kd, -SHIFT);
    double rr = dadd(ylogx, -kd);
    uint64_t t = BB_TAB_EXP2[ki % 32];
    t += ki << (52 - 5);
    double s = u2d(t);
    double C0 = u2d(BB_TAB_EXP2P[0]), C1 = u2d(BB_TAB_EXP2P[1]), C2 = u2d(BB_TAB_EXP2P[2]);
    double zz = dfma(C0, rr, C1, fused);
    double rr2 = dmul(rr, rr);
    double e = dfma(C2, rr, 1.0, fused);
    e = dfma(zz, rr2, e, fused);
    e = dmul(e, s);
    return (float)e;
}

}  // namespace bbpow
