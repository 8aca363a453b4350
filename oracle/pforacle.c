// oracle/pforacle.c — TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
//
// C helpers for the CPU oracle: vector wrappers around the bit-reproducible math contract
// (pathfinder_b200/csrc/pf_math.h) and the random-number contract (pf_rng.h), so that the
// NumPy oracle (oracle/pf_oracle.py, oracle/psis.py) can evaluate *exactly* the floating-point
// operations the device evaluates.  Built by oracle/Makefile with -ffp-contract=off.
#include <stddef.h>
#include "../pathfinder_b200/csrc/pf_rng.h"

void pfo_exp_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_exp(x[i]); }
void pfo_log_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_log(x[i]); }
void pfo_log1p_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_log1p(x[i]); }
void pfo_expm1_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_expm1(x[i]); }

void pfo_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                uint64_t* out2) {
    pf_philox4x32_10(c0, c1, c2, c3, k0, k1, &out2[0], &out2[1]);
}

// Philox4x32 with an arbitrary round count (7 = the normal stream of contract v2).
void pfo_philox_r(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                  uint32_t* out4) {
    uint32_t c[4] = {c0, c1, c2, c3};
    for (int r = 0; r < rounds; ++r) {  // one round at a time through the contract's own round function
        uint32_t o[4];
        pf_philox4x32(1, c[0], c[1], c[2], c[3], k0, k1, o);
        c[0] = o[0]; c[1] = o[1]; c[2] = o[2]; c[3] = o[3];
        k0 += PF_PHILOX_W0; k1 += PF_PHILOX_W1;
    }
    out4[0] = c[0]; out4[1] = c[1]; out4[2] = c[2]; out4[3] = c[3];
}

// the contract call of the normal stream, for tests
void pfo_bits4(uint32_t row_pair, uint32_t stream, uint32_t draw_pair, uint64_t seed, uint32_t call, uint32_t* out4) {
    pf_bits4(row_pair, stream, draw_pair, (uint32_t)seed, (uint32_t)(seed >> 32), call, out4);
}

// u[n x K] column-major standard normals of the engine contract for one (path, iteration) seed.
void pfo_normals(uint64_t seed, int n, int K, double* u) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int k = 0; k < K; ++k) {
        for (int i = 0; i < n; i += 2) {
            double z0, z1;
            pf_normal_pair((uint32_t)(i >> 1), (uint32_t)k, k0, k1, PF_ZIG_XK_INIT, PF_ZIG_F_INIT,
                           &z0, &z1);
            u[(size_t)k * n + i] = z0;
            if (i + 1 < n) u[(size_t)k * n + i + 1] = z1;
        }
    }
}

// 64 random bits per resample draw t = 0..m-1.
void pfo_resample_bits(uint64_t seed, size_t m, uint64_t* out) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (size_t t = 0; t < m; ++t) out[t] = pf_resample_bits(t, k0, k1);
}

uint64_t pfo_mulhi64(uint64_t r, uint64_t z) { return pf_mulhi64(r, z); }

void pfo_zig_tables(uint64_t* xk, double* f) {
    for (int i = 0; i < PF_ZIG_LAYERS; ++i) xk[i] = PF_ZIG_XK_INIT[i];
    for (int i = 0; i <= PF_ZIG_LAYERS; ++i) f[i] = PF_ZIG_F_INIT[i];
}

// one element of the contract, through the element-wise definition (pf_normal_elem)
double pfo_normal_elem(uint64_t seed, uint32_t row, uint32_t draw) {
    return pf_normal_elem(row, draw, (uint32_t)seed, (uint32_t)(seed >> 32), PF_ZIG_XK_INIT, PF_ZIG_F_INIT);
}

// Streaming statistics of the normal stream over `nseeds` seeds x (n x K) variates each, without
// storing them: out = { count, sum, sum2, sum3, sum4, #|z| > thr[0..nthr), slow-path count },
// hist[nbins] over equal-probability bins given by their upper edges `edges[nbins-1]` (ascending).
void pfo_normal_stats(uint64_t seed0, int nseeds, int n, int K, int nthr, const double* thr, int nbins,
                      const double* edges, double* out, uint64_t* hist) {
    double cnt = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, slow = 0;
    for (int a = 0; a < nthr; ++a) out[5 + a] = 0.0;
    for (int b = 0; b < nbins; ++b) hist[b] = 0;
    for (int sd = 0; sd < nseeds; ++sd) {
        const uint64_t seed = seed0 + (uint64_t)sd * 0x9E3779B97F4A7C15ULL;
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        for (int kp = 0; kp < K; kp += 16) {
            for (int g = 0; g < 8 && kp + g < K; ++g) {
                for (int rp = 0; rp < (n + 1) / 2; ++rp) {
                    uint32_t o[4];
                    pf_bits4((uint32_t)rp, 0u, pf_draw_pair((uint32_t)(kp + g)), k0, k1, 0u, o);
                    for (int wd = 0; wd < 4; ++wd) {
                        const int k = kp + g + 8 * (wd >> 1), i = 2 * rp + (wd & 1);
                        if (k >= K || i >= n) continue;
                        double z;
                        if (!pf_zig_fast32(o[wd], PF_ZIG_XK_INIT, &z)) {
                            z = pf_zig_slow32(o[wd], (uint32_t)rp, pf_draw_pair((uint32_t)k), 1u + (uint32_t)wd, k0, k1,
                                              PF_ZIG_XK_INIT, PF_ZIG_F_INIT);
                            slow += 1;
                        }
                        const double z2 = z * z;
                        cnt += 1; s1 += z; s2 += z2; s3 += z2 * z; s4 += z2 * z2;
                        const double az = z < 0 ? -z : z;
                        for (int a = 0; a < nthr; ++a) out[5 + a] += az > thr[a];
                        int lo = 0, hi = nbins - 1;  // first bin whose upper edge exceeds z
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (z < edges[mid]) hi = mid; else lo = mid + 1;
                        }
                        hist[lo] += 1;
                    }
                }
            }
        }
    }
    out[0] = cnt; out[1] = s1; out[2] = s2; out[3] = s3; out[4] = s4;
    out[5 + nthr] = slow;
}
