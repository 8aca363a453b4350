// pf_rng.h — the engine's random-number contract (host + device, same source).
//
// The reference draws u ~ N(0, I) with Julia's Random.randn! on an RNG reseeded per
// iteration from a UInt64 seed (reference: src/elbo.jl:2-5, src/mvnormal.jl:30).  Julia's
// generator cannot be reproduced outside Julia, so the engine fixes its own contract:
//
//   * bits:    Philox4x32-10 (Salmon et al., SC'11) in counter mode with a FIXED key
//              (PF_KEY0, PF_KEY1) and the counter
//                  c0 = row_pair | stream << 28,  c1 = draw,
//                  c2 = seed_lo + call,           c3 = seed_hi,
//              seed = the per-(path, iteration) UInt64 seed.  A fixed key makes the ten round
//              keys compile-time immediates (no per-call key schedule on the GPU's half-rate
//              integer ALU); distinct seeds select disjoint counter sets.
//   * normals: 1024-layer ziggurat (Marsaglia & Tsang 2000) on 64 bits per variate
//              (bit 63 sign, bits 53-62 layer, bits 0-51 a 52-bit mantissa j; x = j 2^-52 x_layer).
//              One Philox call (stream 0, call 0) yields the fast-path words of elements
//              (2*row_pair, draw) and (2*row_pair + 1, draw); an element that leaves the fast
//              path (0.43 %) continues on its private stream (1 + (row & 1), call = 0, 1, ...).
//   * uniforms for resampling: stream 3 (see pf_resample_bits).
//
// The same function bodies are compiled by gcc (oracle helpers, tests) and nvcc (kernels).
#pragma once
#include "pf_math.h"
#include "pf_zig_tables.h"

#define PF_KEY0 0xA4093822u
#define PF_KEY1 0x299F31D0u
#define PF_PHILOX_M0 0xD2511F53u
#define PF_PHILOX_M1 0xCD9E8D57u
#define PF_PHILOX_W0 0x9E3779B9u
#define PF_PHILOX_W1 0xBB67AE85u

PF_HD void pf_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
        : "=r"(*lo), "=r"(*hi) : "r"(a), "r"(b));  // one IMAD.WIDE.U32
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

// Contract call: fixed key, (row_pair, stream, draw, seed, call) in the counter.
PF_HD void pf_bits(uint32_t row_pair, uint32_t stream, uint32_t draw, uint32_t s_lo, uint32_t s_hi,
                   uint32_t call, uint64_t* a, uint64_t* b) {
    pf_philox4x32_10(row_pair | (stream << 28), draw, s_lo + call, s_hi, PF_KEY0, PF_KEY1, a, b);
}

// Uniform in the open interval (0,1) from the top 53 bits: (j + 0.5) * 2^-53.
PF_HD double pf_u01(uint64_t bits) {
    return ((double)(bits >> 11) + 0.5) * 1.1102230246251565e-16;
}

// ---- ziggurat word layout -------------------------------------------------------------------
//   bit 63  sign        bits 53-62  layer i (10 bits)      bits 0-51  mantissa j       (52 unused)
// x = j * w[i] (w[i] = x_i 2^-52).  The layout is chosen for the GPU: the mantissa is the low
// word plus 20 bits of the high word, so (hi & 0xFFFFF) | 0x3FF00000 : lo is the double
// m = 1 + j 2^-52 with one LOP3, and x = fma(m, x_i, -x_i) is one DFMA — exactly j * w[i] rounded
// once, because the product m * x_i is formed exactly inside the fma.  The sign is bit 31 of the
// high word, so it is XORed into x without a shift.
#define PF_ZIG_MANT_MASK 0x000FFFFFFFFFFFFFULL
#define PF_ZIG_ONE_BITS 0x3FF0000000000000ULL

// High word of the fast-accept threshold: the fast test compares only the top 20 mantissa
// bits, (j >> 32) < (kq >> 32) — conservative by < 2^-20; the exact test j < kq is the first
// thing the slow path does, so the variate is the same as with an exact fast test.
PF_HD uint32_t pf_zig_kqh(uint64_t kq) { return 0x3FF00000u | (uint32_t)(kq >> 32); }

// layer edge x_i = 2^52 * w[i] from the bits of w (exponent + 52): exact.
PF_HD double pf_zig_edge(double w) { return pf_u2d(pf_d2u(w) + 0x0340000000000000ULL); }

// x = j * w with the sign of bit 63 (bit-identical on every path).
PF_HD double pf_zig_value(uint64_t bits, double w) {
    double m = pf_u2d((bits & PF_ZIG_MANT_MASK) | PF_ZIG_ONE_BITS);
    double xe = pf_zig_edge(w);
    double x = fma(m, xe, -xe);
    return pf_u2d(pf_d2u(x) ^ (bits & 0x8000000000000000ULL));
}

// Ziggurat fast path.  Returns 1 and writes *z when the variate is accepted without
// evaluating exp/log (99.57 % of calls); *z is written (as if accepted) in either case.
PF_HD int pf_zig_fast(uint64_t bits, const pf_zig_kw_t* kw, double* z) {
    pf_zig_kw_t e = kw[(bits >> 53) & (PF_ZIG_LAYERS - 1)];
    *z = pf_zig_value(bits, e.w);
    uint32_t mh = 0x3FF00000u | ((uint32_t)(bits >> 32) & 0xFFFFFu);
    return mh < pf_zig_kqh(e.kq);
}

// Full ziggurat continuation for an element whose first word `bits` failed the fast test.
// Consumes Philox calls (row_pair, draw, call = 0,1,..., stream) of that element.
PF_HD double pf_zig_slow(uint64_t bits, uint32_t row_pair, uint32_t draw, uint32_t stream,
                         uint32_t k0, uint32_t k1, const pf_zig_kw_t* kw, const double* ftab) {
    uint32_t call = 0;
    for (;;) {
        uint32_t i = (uint32_t)(bits >> 53) & (PF_ZIG_LAYERS - 1);
        uint64_t j = bits & PF_ZIG_MANT_MASK;
        pf_zig_kw_t e = kw[i];
        double z = pf_zig_value(bits, e.w);
        if (j < e.kq) return z;  // exact core test (the fast test is conservative)
        int neg = (int)(bits >> 63);
        uint64_t a, b;
        if (i == 0) {
            // tail beyond r: Marsaglia's exponential-rejection method
            for (;;) {
                pf_bits(row_pair, stream, draw, k0, k1, call++, &a, &b);
                double xt = -pf_log(pf_u01(a)) / PF_ZIG_R;
                double yt = -pf_log(pf_u01(b));
                if (yt + yt > xt * xt) {
                    double x = PF_ZIG_R + xt;
                    return neg ? -x : x;
                }
            }
        }
        double x = neg ? -z : z;
        pf_bits(row_pair, stream, draw, k0, k1, call++, &a, &b);
        double f_lo = ftab[i], f_hi = ftab[i + 1];
        double y = fma(pf_u01(a), f_hi - f_lo, f_lo);
        if (y < pf_exp(-0.5 * x * x)) return z;
        bits = b;  // rejected: b is a fresh first word
    }
}

// The two standard normals of elements (2*row_pair, draw) and (2*row_pair + 1, draw).
PF_HD_NOINLINE void pf_normal_pair(uint32_t row_pair, uint32_t draw, uint32_t k0, uint32_t k1,
                          const pf_zig_kw_t* kw, const double* ftab, double* z0, double* z1) {
    uint64_t a, b;
    pf_bits(row_pair, 0u, draw, k0, k1, 0u, &a, &b);
    if (!pf_zig_fast(a, kw, z0)) *z0 = pf_zig_slow(a, row_pair, draw, 1u, k0, k1, kw, ftab);
    if (!pf_zig_fast(b, kw, z1)) *z1 = pf_zig_slow(b, row_pair, draw, 2u, k0, k1, kw, ftab);
}

// Slow-path continuation of element (row, draw) from scratch: recomputes the element's first
// word (so that any thread can finish any element; used by the warp-balanced deferred slow
// path of K3).  Must only be called for elements whose first word failed pf_zig_fast.
PF_HD_NOINLINE double pf_normal_finish_slow(uint32_t row, uint32_t draw, uint32_t k0, uint32_t k1,
                                            const pf_zig_kw_t* kw, const double* ftab) {
    uint64_t a, b;
    pf_bits(row >> 1, 0u, draw, k0, k1, 0u, &a, &b);
    return pf_zig_slow((row & 1u) ? b : a, row >> 1, draw, 1u + (row & 1u), k0, k1, kw, ftab);
}

// 64 random bits for resample draw t (two per Philox call), stream 3.
PF_HD uint64_t pf_resample_bits(uint64_t t, uint32_t k0, uint32_t k1) {
    uint64_t a, b;
    uint64_t q = t >> 1;
    pf_bits((uint32_t)q & 0x0FFFFFFFu, 3u, (uint32_t)(q >> 28), k0, k1, 0u, &a, &b);
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
