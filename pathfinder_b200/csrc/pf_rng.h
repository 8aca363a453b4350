// pf_rng.h — the engine's random-number contract, version 2 (host + device, same source).
//
// The reference draws u ~ N(0, I) with Julia's Random.randn! on an RNG reseeded per
// iteration from a UInt64 seed (reference: src/elbo.jl:2-5, src/mvnormal.jl:30).  Julia's
// generator cannot be reproduced outside Julia (parity unpinned), so the engine fixes
// its own contract.  Version 2 is shaped by what the sampling kernel K3 measured on B200: the
// 32 x 32 -> 64 multiplies of Philox contend with the FP64 pipe that carries the Woodbury
// products, so the contract spends as few of them per variate as a Crush-resistant generator allows:
//
//   * bits:    Philox4x32-7 (Salmon et al., SC'11: seven rounds is the fewest that passes BigCrush;
//              ten is their safety-margin default, kept here for the resampling stream) in counter
//              mode with a FIXED key (PF_KEY0, PF_KEY1) and the counter
//                  c0 = row_pair | stream << 28,  c1 = draw_pair,
//                  c2 = seed_lo + call,           c3 = seed_hi,
//              seed = the per-(path, iteration) UInt64 seed.  A fixed key makes the round keys
//              compile-time immediates; distinct seeds select disjoint counter sets.
//   * one call = FOUR variates: the 2 x 2 patch  rows {2 row_pair, 2 row_pair + 1}  x
//              draws {k, k + 8}  (k mod 16 < 8) — exactly what one lane of K3 owns in an 8-row
//              block of its two 8-draw sets.  For element (row i, draw k):
//                  row_pair = i >> 1,  draw_pair = (k >> 4) * 8 + (k & 7),
//                  word     = 2 * ((k >> 3) & 1) + (i & 1)      (which of the four 32-bit outputs).
//   * normals: 1024-layer ziggurat (Marsaglia & Tsang 2000) on 32 bits per variate
//              (bit 31 sign, bits 21-30 layer, bit 20 unused, bits 0-19 a 20-bit mantissa j;
//              x = j 2^-20 x_layer: a 2^-20 grid inside each layer, 2^31 distinct values).  An
//              element that leaves the table-only fast path (0.43 %) continues on its private
//              stream (1 + word, call = 0, 1, ...).
//   * uniforms for resampling: stream 3 of Philox4x32-10 (see pf_resample_bits).
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
#define PF_NORMAL_ROUNDS 7

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

// Philox4x32-R: `rounds` is a literal at every call site, so the loop unrolls and the round keys
// fold into immediates.
PF_HD void pf_philox4x32(const int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                         uint32_t k0, uint32_t k1, uint32_t* o) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < rounds; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        pf_mulhilo32(PF_PHILOX_M0, c0, &hi0, &lo0);
        pf_mulhilo32(PF_PHILOX_M1, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PF_PHILOX_W0;
        k1 += PF_PHILOX_W1;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// Philox4x32-10.  a = out[0] | out[1] << 32, b = out[2] | out[3] << 32.
PF_HD void pf_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                            uint32_t k0, uint32_t k1, uint64_t* a, uint64_t* b) {
    uint32_t o[4];
    pf_philox4x32(10, c0, c1, c2, c3, k0, k1, o);
    *a = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    *b = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
}

// Contract call of the resampling stream (ten rounds): fixed key, (index, stream, hi, seed, call).
PF_HD void pf_bits(uint32_t row_pair, uint32_t stream, uint32_t draw, uint32_t s_lo, uint32_t s_hi,
                   uint32_t call, uint64_t* a, uint64_t* b) {
    pf_philox4x32_10(row_pair | (stream << 28), draw, s_lo + call, s_hi, PF_KEY0, PF_KEY1, a, b);
}

// Contract call of the normal stream (seven rounds): the four 32-bit words of one 2 x 2 patch.
PF_HD void pf_bits4(uint32_t row_pair, uint32_t stream, uint32_t draw_pair, uint32_t s_lo, uint32_t s_hi,
                    uint32_t call, uint32_t* o) {
    pf_philox4x32(PF_NORMAL_ROUNDS, row_pair | (stream << 28), draw_pair, s_lo + call, s_hi, PF_KEY0, PF_KEY1, o);
}

// (draw_pair, half) of draw k: the calls are shared by draws k and k + 8 (k mod 16 < 8).
PF_HD uint32_t pf_draw_pair(uint32_t k) { return ((k >> 4) << 3) | (k & 7u); }
PF_HD uint32_t pf_draw_half(uint32_t k) { return (k >> 3) & 1u; }

// Uniform in the open interval (0,1) from the top 53 bits: (j + 0.5) * 2^-53.
PF_HD double pf_u01(uint64_t bits) {
    return ((double)(bits >> 11) + 0.5) * 1.1102230246251565e-16;
}

// ---- ziggurat word layout -------------------------------------------------------------------
//   bit 31  sign        bits 21-30  layer i (10 bits)      bit 20  unused      bits 0-19  mantissa j
// Table entry PF_ZIG_XK[i] (8 bytes, one LDS.64 on the GPU): the double x_i (layer edge) whose low
// 20 mantissa bits carry the fast-accept threshold kq_i, so that entry-as-double IS the edge the
// contract uses (x_i perturbed by < 2^-32 relative — the layers keep equal areas to that accuracy,
// far below the 2^-20 grid) and entry & 0xFFFFF is kq_i = floor(x_{i+1} / x_i * 2^20) (computed
// against the largest double the packing can produce, so j < kq_i implies x < x_{i+1} exactly).
// x = j 2^-20 x_i is formed as fma(m, x_i, -x_i) with m = 1 + j 2^-20 = the double whose high word
// is (w & 0xFFFFF) | 0x3FF00000 and whose low word is 0: one LOP3 and one DFMA, rounded once.
// The sign is bit 31 of the word, XORed into the high word of x without a shift.
#define PF_ZIG_MANT_MASK32 0x000FFFFFu

PF_HD double pf_zig_value32(uint32_t w, uint64_t xk) {
    double m = pf_u2d((uint64_t)(0x3FF00000u | (w & PF_ZIG_MANT_MASK32)) << 32);
    double xe = pf_u2d(xk);
    double x = fma(m, xe, -xe);
    return pf_u2d(pf_d2u(x) ^ ((uint64_t)(w & 0x80000000u) << 32));
}

// Ziggurat fast path.  Returns 1 when the variate is accepted without evaluating exp/log
// (99.57 % of calls); *z is written (as if accepted) in either case.  The test is the exact core
// test (the whole mantissa is compared), so the slow path starts at the wedge / tail.
PF_HD int pf_zig_fast32(uint32_t w, const uint64_t* xk, double* z) {
    uint64_t e = xk[(w >> 21) & (PF_ZIG_LAYERS - 1)];
    *z = pf_zig_value32(w, e);
    return (w & PF_ZIG_MANT_MASK32) < ((uint32_t)e & PF_ZIG_MANT_MASK32);
}

// Full ziggurat continuation for an element whose first word `w` failed the fast test.
// Consumes Philox calls (row_pair, draw_pair, call = 0,1,..., stream) of that element.
PF_HD double pf_zig_slow32(uint32_t w, uint32_t row_pair, uint32_t draw_pair, uint32_t stream,
                           uint32_t k0, uint32_t k1, const uint64_t* xk, const double* ftab) {
    uint32_t call = 0;
    for (;;) {
        uint32_t i = (w >> 21) & (PF_ZIG_LAYERS - 1);
        uint64_t e = xk[i];
        double z = pf_zig_value32(w, e);
        if ((w & PF_ZIG_MANT_MASK32) < ((uint32_t)e & PF_ZIG_MANT_MASK32)) return z;
        int neg = (int)(w >> 31);
        uint32_t o[4];
        if (i == 0) {
            // tail beyond r = x_1: Marsaglia's exponential-rejection method
            const double r = pf_u2d(xk[1]);
            for (;;) {
                pf_bits4(row_pair, stream, draw_pair, k0, k1, call++, o);
                uint64_t a = (uint64_t)o[0] | ((uint64_t)o[1] << 32), b = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
                double xt = -pf_log(pf_u01(a)) / r;
                double yt = -pf_log(pf_u01(b));
                if (yt + yt > xt * xt) {
                    double x = r + xt;
                    return neg ? -x : x;
                }
            }
        }
        double x = neg ? -z : z;
        pf_bits4(row_pair, stream, draw_pair, k0, k1, call++, o);
        uint64_t a = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
        double f_lo = ftab[i], f_hi = ftab[i + 1];
        double y = fma(pf_u01(a), f_hi - f_lo, f_lo);
        if (y < pf_exp(-0.5 * x * x)) return z;
        w = o[2];  // rejected: a fresh first word
    }
}

// The standard normal of element (row i, draw k) — the contract, element by element.
PF_HD_NOINLINE double pf_normal_elem(uint32_t row, uint32_t draw, uint32_t k0, uint32_t k1,
                                     const uint64_t* xk, const double* ftab) {
    const uint32_t rp = row >> 1, dp = pf_draw_pair(draw), word = 2u * pf_draw_half(draw) + (row & 1u);
    uint32_t o[4];
    pf_bits4(rp, 0u, dp, k0, k1, 0u, o);
    const uint32_t w = word == 0 ? o[0] : (word == 1 ? o[1] : (word == 2 ? o[2] : o[3]));
    double z;
    if (pf_zig_fast32(w, xk, &z)) return z;
    return pf_zig_slow32(w, rp, dp, 1u + word, k0, k1, xk, ftab);
}

// The two standard normals of elements (2*row_pair, draw) and (2*row_pair + 1, draw).
PF_HD_NOINLINE void pf_normal_pair(uint32_t row_pair, uint32_t draw, uint32_t k0, uint32_t k1,
                                   const uint64_t* xk, const double* ftab, double* z0, double* z1) {
    const uint32_t dp = pf_draw_pair(draw), h = pf_draw_half(draw);
    uint32_t o[4];
    pf_bits4(row_pair, 0u, dp, k0, k1, 0u, o);
    const uint32_t wa = h ? o[2] : o[0], wb = h ? o[3] : o[1];
    if (!pf_zig_fast32(wa, xk, z0)) *z0 = pf_zig_slow32(wa, row_pair, dp, 1u + 2u * h, k0, k1, xk, ftab);
    if (!pf_zig_fast32(wb, xk, z1)) *z1 = pf_zig_slow32(wb, row_pair, dp, 2u + 2u * h, k0, k1, xk, ftab);
}

// Slow-path continuation of element (row, draw) from scratch: recomputes the element's first
// word (so that any thread can finish any element; used by the warp-balanced deferred slow
// path of K3).  Must only be called for elements whose first word failed pf_zig_fast32.
// Returns the variate and the provisional value the fast path computed from the first word — the
// two are equal whenever the wedge test accepts the first candidate.
typedef struct { double z, zprov; } pf_slow_t;
PF_HD_NOINLINE pf_slow_t pf_normal_finish_slow(uint32_t row, uint32_t draw, uint32_t k0, uint32_t k1,
                                               const uint64_t* xk, const double* ftab) {
    const uint32_t rp = row >> 1, dp = pf_draw_pair(draw), word = 2u * pf_draw_half(draw) + (row & 1u);
    uint32_t o[4];
    pf_bits4(rp, 0u, dp, k0, k1, 0u, o);
    const uint32_t w = word == 0 ? o[0] : (word == 1 ? o[1] : (word == 2 ? o[2] : o[3]));
    pf_slow_t r;
    r.zprov = pf_zig_value32(w, xk[(w >> 21) & (PF_ZIG_LAYERS - 1)]);
    r.z = pf_zig_slow32(w, rp, dp, 1u + word, k0, k1, xk, ftab);
    return r;
}

// 64 random bits for resample draw t (two per Philox4x32-10 call), stream 3.
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
