// pf_math.h — bit-reproducible double-precision transcendentals for host and device.
//
// Why this exists: the PSIS / resample stage (reference: src/resample.jl:58-95 plus
// PSIS.jl's psis / fit_gpd) must give *bit-identical* importance weights on the CPU oracle
// and on the GPU so that resample indices can be compared exactly.  libm and libdevice
// disagree in the last ulp, so both sides evaluate the same sequence of IEEE-754 basic
// operations (+ - * / sqrt, all correctly rounded) and explicit fma() calls from this header.
//
// Build contract: every translation unit that includes this header for a *bit-exact* path
// must be compiled with floating-point contraction disabled (nvcc -fmad=false,
// gcc -ffp-contract=off); fused multiply-adds happen only where fma() is written out.
// Accuracy (measured in tests/test_pf_math.py against libm): <= 2 ulp on the tested ranges.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define PF_HD __host__ __device__ __forceinline__
#define PF_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define PF_HD static inline
#define PF_HD_NOINLINE static __attribute__((noinline))
#endif

PF_HD uint64_t pf_d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
PF_HD double pf_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

// 2^k for k in [-1022, 1023], built from the exponent field.
PF_HD double pf_pow2i(int k) { return pf_u2d((uint64_t)(k + 1023) << 52); }

// exp(x).  Range reduction x = k*ln2 + r, |r| <= ln2/2, Cody-Waite two-constant split with
// fma; degree-13 Taylor polynomial in Horner form (truncation < 2^-57 on the reduced range).
PF_HD double pf_exp(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.1332191019412) return 0.0;
    const double INV_LN2 = 1.4426950408889634074;
    const double LN2_HI = 6.93147180369123816490e-01;  // fdlibm split of ln 2
    const double LN2_LO = 1.90821492927058770002e-10;
    double kd = floor(fma(x, INV_LN2, 0.5));
    int k = (int)kd;
    double r = fma(-kd, LN2_HI, x);
    r = fma(-kd, LN2_LO, r);
    double p = 1.0 / 6227020800.0;             // 1/13!
    p = fma(p, r, 1.0 / 479001600.0);           // 1/12!
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // scale by 2^k in two exact steps so subnormal results round once at the end
    int k1 = k / 2, k2 = k - k1;
    return (p * pf_pow2i(k1)) * pf_pow2i(k2);
}

// log(x), fdlibm e_log.c formulation: x = 2^e * (1+f), sqrt(2)/2 < 1+f < sqrt(2),
// s = f/(2+f), log(1+f) = f - (f^2/2 - s*(f^2/2 + R(s^2))).
PF_HD double pf_log(double x) {
    if (x != x) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                 Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int e = 0;
    uint64_t u = pf_d2u(x);
    if ((u >> 52) == 0) {  // subnormal: scale up by 2^54
        x = x * 18014398509481984.0;
        u = pf_d2u(x);
        e = -54;
    }
    e += (int)(u >> 52) - 1023;
    uint64_t m = u & 0x000FFFFFFFFFFFFFULL;
    // mantissa in [1,2); move to [sqrt2/2, sqrt2)
    if (m >= 0x6A09E667F3BCDULL) {  // 1.m >= sqrt(2)
        u = m | 0x3FE0000000000000ULL;  // 1.m / 2
        e += 1;
    } else {
        u = m | 0x3FF0000000000000ULL;
    }
    double f = pf_u2d(u) - 1.0;
    double dk = (double)e;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    // dk*ln2_hi + (f - (hfsq - (s*(hfsq+R) + dk*ln2_lo)))
    return fma(dk, LN2_HI, f - (hfsq - fma(s, hfsq + R, dk * LN2_LO)));
}

// log1p(x) = log(1+x) with the Kahan/HP-15C correction log(u) * x / (u - 1), u = fl(1+x).
PF_HD double pf_log1p(double x) {
    double u = 1.0 + x;
    if (u == 1.0) return x;
    if (u == INFINITY) return u;
    return pf_log(u) * (x / (u - 1.0));
}

// expm1(x) = exp(x) - 1 with the Kahan correction (u - 1) * x / log(u), u = exp(x).
PF_HD double pf_expm1(double x) {
    double u = pf_exp(x);
    if (u == 1.0) return x;
    double um1 = u - 1.0;
    if (um1 == -1.0) return -1.0;
    if (u == INFINITY) return u;
    return um1 * (x / pf_log(u));
}
