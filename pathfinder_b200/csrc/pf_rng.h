// pf_rng.h — the engine's random-number contract (host + device, same source).
//
// The reference draws u ~ N(0, I) with Julia's Random.randn! on an RNG reseeded per
// iteration from a UInt64 seed (reference: src/elbo.jl:2-5, src/mvnormal.jl:30).  Julia's
// generator cannot be reproduced outside Julia, so the engine fixes its own contract:
//
//   * bits:    Philox4x32-10 (Salmon et al., SC'11), key = the per-(path, iteration) UInt64
//              seed, counter = (row_pair, draw, call, stream).
//   * normals: 256-layer ziggurat (Marsaglia & Tsang 2000) on 64 bits per variate
//              (bits 0-7 layer, bit 8 sign, bits 12-63 a 52 bit mantissa j; x = j * 2^-52 * x_layer).  One Philox call at
//              (row_pair = i/2, draw = k, call = 0, stream = 0) yields the fast-path words of
//              elements (2*(i/2), k) and (2*(i/2)+1, k); an element that leaves the fast path
//              (~1.2 %) continues on its private stream (stream = 1 + (i & 1), call = 0,1,...).
//   * uniforms for resampling: stream = 3 (see pf_resample_bits).
//
// The same function bodies are compiled by gcc (oracle helpers, tests) and nvcc (kernels).
#pragma once
#include "pf_math.h"
#include "pf_zig_tables.h"

#define PF_PHILOX_M0 0xD2511F53u
#define PF_PHILOX_M1 0xCD9E8D57u
#define PF_PHILOX_W0 0x9E3779B9u
#define PF_PHILOX_W1 0xBB67AE85u

PF_HD void pf_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *lo = (uint32_t)p;
    *hi = (uint32_t)(p >> 32);
#endif
}

// Philox4x32-10.  out[0..3]; a = out[0] | out[1] << 32, b = out[2] | out[3] << 32.
PF_HD void pf_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                            uint32_t k0, uint32_t k1, uint64_t* a, uint64_t* b) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        pf_mulhilo32(PF_PHILOX_M0, c0, &hi0, &lo0);
        pf_mulhilo32(PF_PHILOX_M1, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PF_PHILOX_W0;
        k1 += PF_PHILOX_W1;
    }
    *a = (uint64_t)c0 | ((uint64_t)c1 << 32);
    *b = (uint64_t)c2 | ((uint64_t)c3 << 32);
}

// Uniform in the open interval (0,1) from the top 53 bits: (j + 0.5) * 2^-53.
PF_HD double pf_u01(uint64_t bits) {
    return ((double)(bits >> 11) + 0.5) * 1.1102230246251565e-16;
}

// Exact conversion of a 52-bit integer to double without an int64->fp64 convert instruction:
// (2^52 + j) - 2^52.
PF_HD double pf_mant52(uint64_t j) {
    return pf_u2d(0x4330000000000000ULL | j) - 4503599627370496.0;
}

// Ziggurat fast path.  Returns 1 and writes *z when the variate is accepted without
// evaluating exp/log (98.8 % of calls).
PF_HD int pf_zig_fast(uint64_t bits, const pf_zig_kw_t* kw, double* z) {
    uint32_t i = (uint32_t)bits & 255u;
    uint64_t j = bits >> 12;
    pf_zig_kw_t e = kw[i];
    double x = pf_mant52(j) * e.w;
    *z = ((bits >> 8) & 1u) ? -x : x;
    return j < e.kq;
}

// Full ziggurat continuation for an element whose first word `bits` failed the fast test.
// Consumes Philox calls (row_pair, draw, call = 0,1,..., stream) of that element.
PF_HD double pf_zig_slow(uint64_t bits, uint32_t row_pair, uint32_t draw, uint32_t stream,
                         uint32_t k0, uint32_t k1, const pf_zig_kw_t* kw, const double* ftab) {
    uint32_t call = 0;
    for (;;) {
        uint32_t i = (uint32_t)bits & 255u;
        uint64_t j = bits >> 12;
        int neg = (int)((bits >> 8) & 1u);
        uint64_t a, b;
        if (i == 0) {
            // tail beyond r: Marsaglia's exponential-rejection method
            for (;;) {
                pf_philox4x32_10(row_pair, draw, call++, stream, k0, k1, &a, &b);
                double xt = -pf_log(pf_u01(a)) / PF_ZIG_R;
                double yt = -pf_log(pf_u01(b));
                if (yt + yt > xt * xt) {
                    double x = PF_ZIG_R + xt;
                    return neg ? -x : x;
                }
            }
        }
        double x = pf_mant52(j) * kw[i].w;
        pf_philox4x32_10(row_pair, draw, call++, stream, k0, k1, &a, &b);
        double f_lo = ftab[i], f_hi = ftab[i + 1];
        double y = fma(pf_u01(a), f_hi - f_lo, f_lo);
        if (y < pf_exp(-0.5 * x * x)) return neg ? -x : x;
        // rejected: b is a fresh first word
        bits = b;
        double z;
        if (pf_zig_fast(bits, kw, &z)) return z;
    }
}

// The two standard normals of elements (2*row_pair, draw) and (2*row_pair + 1, draw).
PF_HD_NOINLINE void pf_normal_pair(uint32_t row_pair, uint32_t draw, uint32_t k0, uint32_t k1,
                          const pf_zig_kw_t* kw, const double* ftab, double* z0, double* z1) {
    uint64_t a, b;
    pf_philox4x32_10(row_pair, draw, 0u, 0u, k0, k1, &a, &b);
    if (!pf_zig_fast(a, kw, z0)) *z0 = pf_zig_slow(a, row_pair, draw, 1u, k0, k1, kw, ftab);
    if (!pf_zig_fast(b, kw, z1)) *z1 = pf_zig_slow(b, row_pair, draw, 2u, k0, k1, kw, ftab);
}

// Slow-path continuation of element (row, draw) from scratch: recomputes the element's first
// word (so that any thread can finish any element; used by the warp-balanced deferred slow
// path of K3).  Must only be called for elements whose first word failed pf_zig_fast.
PF_HD_NOINLINE double pf_normal_finish_slow(uint32_t row, uint32_t draw, uint32_t k0, uint32_t k1,
                                            const pf_zig_kw_t* kw, const double* ftab) {
    uint64_t a, b;
    pf_philox4x32_10(row >> 1, draw, 0u, 0u, k0, k1, &a, &b);
    return pf_zig_slow((row & 1u) ? b : a, row >> 1, draw, 1u + (row & 1u), k0, k1, kw, ftab);
}

// 64 random bits for resample draw t (two per Philox call), stream 3.
PF_HD uint64_t pf_resample_bits(uint64_t t, uint32_t k0, uint32_t k1) {
    uint64_t a, b;
    uint64_t q = t >> 1;
    pf_philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), 0u, 3u, k0, k1, &a, &b);
    return (t & 1) ? b : a;
}

// floor(r * z / 2^64) for r, z < 2^64 : maps 64 random bits onto [0, z).
PF_HD uint64_t pf_mulhi64(uint64_t r, uint64_t z) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(r, z);
#else
    return (uint64_t)(((unsigned __int128)r * (unsigned __int128)z) >> 64);
#endif
}
