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

// u[n x K] column-major standard normals of the engine contract for one (path, iteration) seed.
void pfo_normals(uint64_t seed, int n, int K, double* u) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int k = 0; k < K; ++k) {
        for (int i = 0; i < n; i += 2) {
            double z0, z1;
            pf_normal_pair((uint32_t)(i >> 1), (uint32_t)k, k0, k1, PF_ZIG_KW_INIT, PF_ZIG_F_INIT,
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

void pfo_zig_tables(uint64_t* kq, double* w, double* f) {
    for (int i = 0; i < PF_ZIG_LAYERS; ++i) { kq[i] = PF_ZIG_KW_INIT[i].kq; w[i] = PF_ZIG_KW_INIT[i].w; }
    for (int i = 0; i <= PF_ZIG_LAYERS; ++i) f[i] = PF_ZIG_F_INIT[i];
}
