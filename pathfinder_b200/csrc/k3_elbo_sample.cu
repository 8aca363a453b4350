// K3 elbo_sample_fused (and K5 materialize_best, same kernel with MATERIALIZE = true).
//
// Replaces rand_and_logpdf + the per-draw part of elbo_and_samples:
//   u ~ N(0, I_n)            (reference: src/mvnormal.jl:30; here the engine's Philox/ziggurat
//                             contract of pf_rng.h, or host-supplied normals in parity mode)
//   |u|^2                    (src/mvnormal.jl:31)
//   x = L u + mu, L = U' Q diag(Vc', I)       (src/mvnormal.jl:32-33 -> src/woodbury.jl:136-143)
//       Q applied in compact-WY form  Q = I - Vh T Vh'  (what LAPACK dgemqrt does)
//   logq = -(n log 2pi + logdet + |u|^2) / 2  (src/mvnormal.jl:36)
//   logp = log pi(x) for the registered model  (src/elbo.jl:15)
//
// Mapping: one thread = one Monte-Carlo draw; a CTA = PFB_K3_THREADS draws of one unit.  All
// lanes of a warp read the same factor-record row, so shared-memory reads are broadcasts.
// The unit's factor record (n rows of {Vh[i][0..KP), sqrt(alpha_i), mu_i}) is streamed through
// a ring of shared-memory stages by 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx),
// twice: pass 0 accumulates w = Vh' u~ (KP dot products per draw, in registers), pass 1
// regenerates u from the counter-based RNG and forms x = sqrt(alpha) .* (u~ - Vh (T w)) + mu.
// Nothing but logp/logq (16 B per draw) is written in lean mode; MATERIALIZE also writes x.
#include "pfb_common.cuh"
#include "pf_rng.h"

// One translation unit per padded reflector count (compiled in parallel by the Makefile):
//   -DPFB_K3_KP=12 -DPFB_K3_ENTRY=pfb_launch_k3_kp12   etc.
#ifndef PFB_K3_KP
#define PFB_K3_KP 12
#define PFB_K3_ENTRY pfb_launch_k3_kp12
#endif

#define PFB_K3_THREADS 256
#ifndef PFB_K3_MINBLOCKS
#define PFB_K3_MINBLOCKS 2
#endif
#define PFB_K3_RC 128    // record rows per TMA stage
#define PFB_K3_NSTAGE 3

struct pfb_model_params {
    const double* p0;  // DIAGNORMAL: mean[n]
    const double* p1;  // DIAGNORMAL: 1/sd[n]
    double c0;         // DIAGNORMAL: -sum(log sd) - n/2 log(2 pi)
};

template <int MODEL>
struct pfb_model_acc {
    double a, b;
    __device__ __forceinline__ void init() { a = 0.0; b = 0.0; }
    __device__ __forceinline__ void add(int i, double x, const pfb_model_params& mp) {
        if (MODEL == PFB_MODEL_ISONORMAL) {
            a = fma(x, x, a);
        } else if (MODEL == PFB_MODEL_FUNNEL) {
            if (i == 0) b = x; else a = fma(x, x, a);
        } else if (MODEL == PFB_MODEL_DIAGNORMAL) {
            double z = (x - __ldg(mp.p0 + i)) * __ldg(mp.p1 + i);
            a = fma(z, z, a);
        }
    }
    __device__ __forceinline__ double finish(int n, const pfb_model_params& mp) const {
        if (MODEL == PFB_MODEL_ISONORMAL) return a / -2.0;
        if (MODEL == PFB_MODEL_FUNNEL) {
            // ((tau/3)^2 + (n-1) tau + exp(-tau) * sum beta^2) / -2
            double t3 = b / 3.0;
            return (fma(t3, t3, (double)(n - 1) * b) + exp(-b) * a) / -2.0;
        }
        if (MODEL == PFB_MODEL_DIAGNORMAL) return fma(a, -0.5, mp.c0);
        return NAN;
    }
};

// Deferred, warp-balanced ziggurat slow path.  Elements whose first word fails the fast test
// (1.5 %) are only *recorded* in a per-lane bit mask while the chunk's main loop runs with
// z = 0 for them; at the end of the chunk the warp flattens all pending (lane, row) items into
// a shared list, every lane finishes items round-robin (so the rare exp/log path runs
// convergently and load-balanced instead of stalling 31 lanes for one), and each owner lane
// then folds its results into its accumulators.  Values are identical to pf_normal_pair().
#define PFB_K3_DCAP 128  // list capacity per warp and round

struct pfb_k3_warp_list {
    double z[PFB_K3_DCAP];
    uint16_t row[PFB_K3_DCAP];  // row offset inside the chunk
    uint8_t src[PFB_K3_DCAP];   // owner lane
};

template <int KP, int MODEL>
__global__ void __launch_bounds__(PFB_K3_THREADS, PFB_K3_MINBLOCKS)
pfb_k3_elbo_sample(int n, int K, int tiles_per_unit, const int32_t* __restrict__ unit_list,
                   const double* __restrict__ FR, const double* __restrict__ HDR,
                   const uint64_t* __restrict__ seeds, const double* __restrict__ u_host,
                   pfb_model_params mp, double* __restrict__ logp_out, double* __restrict__ logq_out,
                   double* __restrict__ draws_out) {
    constexpr int RS = KP + 2;
    constexpr int RC = PFB_K3_RC;
    constexpr int NS = PFB_K3_NSTAGE;
    constexpr int NW = PFB_K3_THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sStage = reinterpret_cast<double*>(smem_raw);                  // NS * RC * RS
    double* sT = sStage + NS * RC * RS;                                    // KP*KP
    double* sVc = sT + KP * KP;                                            // KP*KP
    pf_zig_kw_t* sKw = reinterpret_cast<pf_zig_kw_t*>(sVc + KP * KP);      // 256
    double* sF = reinterpret_cast<double*>(sKw + PF_ZIG_LAYERS);           // 257 (+1 pad)
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sF + PF_ZIG_LAYERS + 2);  // NS (+1 pad)
    pfb_k3_warp_list* sList = reinterpret_cast<pfb_k3_warp_list*>(sBar + NS + 1);  // NW

    const int slot = blockIdx.x / tiles_per_unit;
    const int tile = blockIdx.x - slot * tiles_per_unit;
    const int unit = unit_list ? unit_list[slot] : slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kraw = tile * PFB_K3_THREADS + tid;
    const bool active = kraw < K;
    const uint32_t k = (uint32_t)(active ? kraw : K - 1);
    const uint32_t kwarp0 = (uint32_t)(tile * PFB_K3_THREADS + warp * 32);  // draw of lane 0
    if (unit < 0) {  // path without a usable iteration (K5 only)
        if (active) {
            logp_out[(int64_t)slot * K + k] = NAN;
            logq_out[(int64_t)slot * K + k] = NAN;
        }
        return;
    }
    const double* fr = FR + (int64_t)unit * n * RS;
    const double* hdr = HDR + (int64_t)unit * pfb_hs_of(KP);
    const uint64_t seed = seeds[unit];
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const bool HOST_U = (u_host != nullptr);      // parity mode: normals supplied by the host
    const bool MATERIALIZE = (draws_out != nullptr);
    const double* uh = HOST_U ? u_host + ((int64_t)unit * K + k) * n : nullptr;

    const int C = (n + RC - 1) / RC;            // chunks per pass
    const bool resident = (C <= NS);
    const int Q = resident ? C : 2 * C;          // TMA loads in this CTA

    for (int e = tid; e < KP * KP; e += PFB_K3_THREADS) {
        sT[e] = hdr[e];
        sVc[e] = hdr[KP * KP + e];
    }
    for (int e = tid; e < PF_ZIG_LAYERS; e += PFB_K3_THREADS) sKw[e] = PF_ZIG_KW_DEV[e];
    for (int e = tid; e <= PF_ZIG_LAYERS; e += PFB_K3_THREADS) sF[e] = PF_ZIG_F_DEV[e];
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) pfb_mbar_init(&sBar[s], 1);
        pfb_fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int q) {
        const int c = q % C, s = q % NS;
        const int r0 = c * RC;
        const int rows = min(RC, n - r0);
        const uint32_t bytes = (uint32_t)(rows * RS * 8);
        pfb_mbar_expect_tx(&sBar[s], bytes);
        pfb_tma_load_1d(sStage + s * RC * RS, fr + (int64_t)r0 * RS, bytes, &sBar[s]);
    };
    if (tid == 0) {
        for (int q = 0; q < NS && q < Q; ++q) issue(q);
    }
    const double logdet = hdr[PFB_HDR_LOGDET(KP)];
    const bool pd_ok = hdr[PFB_HDR_FLAG(KP)] != 0.0;
    const int H = min(KP, n);  // head rows (get the Vc' multiply)

    double w[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) w[j] = 0.0;
    double unormsq = 0.0;
    pfb_model_acc<MODEL> macc;
    macc.init();
    double* xout = MATERIALIZE ? draws_out + ((int64_t)slot * K + k) * n : nullptr;
    pfb_k3_warp_list& wl = sList[warp];

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            const int q = pass * C + c;
            const int s = resident ? c : (q % NS);
            if (!(resident && pass == 1)) pfb_mbar_wait(&sBar[s], (uint32_t)((q / NS) & 1));
            const double* st = sStage + s * RC * RS;
            const int r0 = c * RC;
            const int r1 = min(n, r0 + RC);
            int ibeg = r0;
            if (c == 0) {
                // ---- head rows 0..H-1: u~ = Vc' u (src/woodbury.jl:139) ----------------------
                double zh[KP];
#pragma unroll
                for (int jj = 0; jj < KP / 2; ++jj) {
                    double z0 = 0.0, z1 = 0.0;
                    if (2 * jj < n) {
                        if (HOST_U) {
                            z0 = uh[2 * jj];
                            if (2 * jj + 1 < n) z1 = uh[2 * jj + 1];
                        } else {
                            pf_normal_pair((uint32_t)jj, k, k0, k1, sKw, sF, &z0, &z1);
                            if (2 * jj + 1 >= n) z1 = 0.0;
                        }
                    }
                    zh[2 * jj] = z0;
                    zh[2 * jj + 1] = z1;
                }
                if (pass == 0) {
#pragma unroll
                    for (int j = 0; j < KP; ++j) unormsq = fma(zh[j], zh[j], unormsq);
                }
                // in place, descending j: zh[j] <- sum_{m<=j} Vc[m][j] zh[m]
#pragma unroll
                for (int j = KP - 1; j >= 0; --j) {
                    double t = 0.0;
#pragma unroll
                    for (int m = 0; m <= j; ++m) t = fma(sVc[m * KP + j], zh[m], t);
                    zh[j] = t;
                }
#pragma unroll
                for (int i = 0; i < KP; ++i) {
                    if (i < H) {
                        const double2* rec = reinterpret_cast<const double2*>(st + i * RS);
                        if (pass == 0) {
#pragma unroll
                            for (int j2 = 0; j2 < KP / 2; ++j2) {
                                double2 v = rec[j2];
                                w[2 * j2] = fma(v.x, zh[i], w[2 * j2]);
                                w[2 * j2 + 1] = fma(v.y, zh[i], w[2 * j2 + 1]);
                            }
                        } else {
                            double z = zh[i];
#pragma unroll
                            for (int j2 = 0; j2 < KP / 2; ++j2) {
                                double2 v = rec[j2];
                                z = fma(-v.x, w[2 * j2], z);
                                z = fma(-v.y, w[2 * j2 + 1], z);
                            }
                            double2 am = rec[KP / 2];
                            double x = fma(am.x, z, am.y);
                            macc.add(i, x, mp);
                            if (MATERIALIZE && active) xout[i] = x;
                        }
                    }
                }
                ibeg = H;
            }
            // ---- body rows, two per Philox call; slow-path elements are deferred --------------
            unsigned long long m0 = 0ull, m1 = 0ull;  // pending rows (offset in chunk) of this lane
#pragma unroll 1
            for (int i = ibeg; i < r1; i += 2) {
                double z0, z1;
                bool ok0 = true, ok1 = true;
                if (HOST_U) {
                    z0 = uh[i];
                    z1 = (i + 1 < n) ? uh[i + 1] : 0.0;
                } else {
                    uint64_t a, b;
                    pf_philox4x32_10((uint32_t)(i >> 1), k, 0u, 0u, k0, k1, &a, &b);
                    ok0 = pf_zig_fast(a, sKw, &z0);
                    ok1 = pf_zig_fast(b, sKw, &z1) || (i + 1 >= n);
                    if (i + 1 >= n) z1 = 0.0;
                    if (!active) { ok0 = true; ok1 = true; }
                    if (!(ok0 && ok1)) {
                        const int o = i - r0;
                        unsigned long long bits = (ok0 ? 0ull : 1ull) | (ok1 ? 0ull : 2ull);
                        if (o < 64) m0 |= bits << o; else m1 |= bits << (o - 64);
                        if (!ok0) z0 = 0.0;
                        if (!ok1) z1 = 0.0;
                    }
                }
                const double2* ra = reinterpret_cast<const double2*>(st + (i - r0) * RS);
                const double2* rb = ra + RS / 2;
                const bool has_b = (i + 1 < r1);
                if (pass == 0) {
                    unormsq = fma(z0, z0, unormsq);
                    unormsq = fma(z1, z1, unormsq);
#pragma unroll
                    for (int j2 = 0; j2 < KP / 2; ++j2) {
                        double2 va = ra[j2];
                        w[2 * j2] = fma(va.x, z0, w[2 * j2]);
                        w[2 * j2 + 1] = fma(va.y, z0, w[2 * j2 + 1]);
                    }
                    if (has_b) {
#pragma unroll
                        for (int j2 = 0; j2 < KP / 2; ++j2) {
                            double2 vb = rb[j2];
                            w[2 * j2] = fma(vb.x, z1, w[2 * j2]);
                            w[2 * j2 + 1] = fma(vb.y, z1, w[2 * j2 + 1]);
                        }
                    }
                } else {
#pragma unroll
                    for (int j2 = 0; j2 < KP / 2; ++j2) {
                        double2 va = ra[j2];
                        z0 = fma(-va.x, w[2 * j2], z0);
                        z0 = fma(-va.y, w[2 * j2 + 1], z0);
                    }
                    double2 am0 = ra[KP / 2];
                    double x0 = fma(am0.x, z0, am0.y);
                    double x1 = 0.0;
                    if (ok0) macc.add(i, x0, mp);
                    if (has_b) {
#pragma unroll
                        for (int j2 = 0; j2 < KP / 2; ++j2) {
                            double2 vb = rb[j2];
                            z1 = fma(-vb.x, w[2 * j2], z1);
                            z1 = fma(-vb.y, w[2 * j2 + 1], z1);
                        }
                        double2 am1 = rb[KP / 2];
                        x1 = fma(am1.x, z1, am1.y);
                        if (ok1) macc.add(i + 1, x1, mp);
                    }
                    if (MATERIALIZE && active) {
                        if (has_b && (n & 1) == 0) {
                            *reinterpret_cast<double2*>(xout + i) = make_double2(x0, x1);
                        } else {
                            xout[i] = x0;
                            if (has_b) xout[i + 1] = x1;
                        }
                    }
                }
            }
            // ---- deferred slow path (warp-synchronous) ------------------------------------------
            if (!HOST_U) {
                const int cnt = __popcll(m0) + __popcll(m1);
                int incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += t;
                }
                const int excl = incl - cnt;
                const int total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll 1
                for (int base = 0; base < total; base += PFB_K3_DCAP) {
                    {
                        unsigned long long a0 = m0, a1 = m1;
                        int flat = excl - base;
                        while (a0 | a1) {
                            int o;
                            if (a0) { o = __ffsll((long long)a0) - 1; a0 &= a0 - 1; }
                            else { o = 64 + __ffsll((long long)a1) - 1; a1 &= a1 - 1; }
                            if (flat >= 0 && flat < PFB_K3_DCAP) {
                                wl.row[flat] = (uint16_t)o;
                                wl.src[flat] = (uint8_t)lane;
                            }
                            ++flat;
                        }
                    }
                    __syncwarp();
                    const int nitems = min(PFB_K3_DCAP, total - base);
                    for (int t = lane; t < nitems; t += 32) {
                        const uint32_t row = (uint32_t)(r0 + wl.row[t]);
                        const uint32_t kd = kwarp0 + wl.src[t];
                        wl.z[t] = pf_normal_finish_slow(row, kd, k0, k1, sKw, sF);
                    }
                    __syncwarp();
                    {
                        unsigned long long a0 = m0, a1 = m1;
                        int flat = excl - base;
                        while (a0 | a1) {
                            int o;
                            if (a0) { o = __ffsll((long long)a0) - 1; a0 &= a0 - 1; }
                            else { o = 64 + __ffsll((long long)a1) - 1; a1 &= a1 - 1; }
                            if (flat >= 0 && flat < PFB_K3_DCAP) {
                                double z = wl.z[flat];
                                const double2* rr = reinterpret_cast<const double2*>(st + o * RS);
                                if (pass == 0) {
                                    unormsq = fma(z, z, unormsq);
#pragma unroll
                                    for (int j2 = 0; j2 < KP / 2; ++j2) {
                                        double2 v = rr[j2];
                                        w[2 * j2] = fma(v.x, z, w[2 * j2]);
                                        w[2 * j2 + 1] = fma(v.y, z, w[2 * j2 + 1]);
                                    }
                                } else {
#pragma unroll
                                    for (int j2 = 0; j2 < KP / 2; ++j2) {
                                        double2 v = rr[j2];
                                        z = fma(-v.x, w[2 * j2], z);
                                        z = fma(-v.y, w[2 * j2 + 1], z);
                                    }
                                    double2 am = rr[KP / 2];
                                    double x = fma(am.x, z, am.y);
                                    macc.add(r0 + o, x, mp);
                                    if (MATERIALIZE) xout[r0 + o] = x;
                                }
                            }
                            ++flat;
                        }
                    }
                    __syncwarp();
                }
            }
            if (!resident) {
                __syncthreads();  // every thread is done with stage s
                if (tid == 0 && q + NS < Q) issue(q + NS);
            }
        }
        if (pass == 0) {
            // w <- T w  (upper triangular, row-major in smem); in place, ascending rows
#pragma unroll
            for (int a = 0; a < KP; ++a) {
                double t = 0.0;
#pragma unroll
                for (int b = a; b < KP; ++b) t = fma(sT[a * KP + b], w[b], t);
                w[a] = t;
            }
        }
    }
    if (active) {
        double logq = (fma((double)n, PFB_LOG2PI, logdet) + unormsq) / -2.0;
        if (!pd_ok) logq = NAN;
        logp_out[(int64_t)slot * K + k] = macc.finish(n, mp);
        logq_out[(int64_t)slot * K + k] = logq;
    }
}

template <int KP, int MODEL>
static cudaError_t launch_k3_m(cudaStream_t st, int n, int K, int nslots, const int32_t* unit_list,
                               const double* FR, const double* HDR, const uint64_t* seeds,
                               const double* u_host, pfb_model_params mp, double* logp, double* logq,
                               double* draws) {
    const int tiles = (K + PFB_K3_THREADS - 1) / PFB_K3_THREADS;
    const size_t smem = (size_t)(PFB_K3_NSTAGE * PFB_K3_RC * (KP + 2) + 2 * KP * KP) * 8 +
                        PF_ZIG_LAYERS * sizeof(pf_zig_kw_t) + (PF_ZIG_LAYERS + 2) * 8 +
                        (PFB_K3_NSTAGE + 1) * 8 + (PFB_K3_THREADS / 32) * sizeof(pfb_k3_warp_list);
    const int64_t grid = (int64_t)nslots * tiles;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidValue;
    auto kern = pfb_k3_elbo_sample<KP, MODEL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)grid, PFB_K3_THREADS, smem, st>>>(n, K, tiles, unit_list, FR, HDR, seeds, u_host, mp,
                                                       logp, logq, draws);
    return cudaGetLastError();
}

template <int KP>
static cudaError_t launch_k3_k(cudaStream_t st, int model, int n, int K, int nslots,
                               const int32_t* unit_list, const double* FR, const double* HDR,
                               const uint64_t* seeds, const double* u_host, pfb_model_params mp,
                               double* logp, double* logq, double* draws) {
    switch (model) {
        case PFB_MODEL_ISONORMAL:
            return launch_k3_m<KP, PFB_MODEL_ISONORMAL>(st, n, K, nslots, unit_list, FR, HDR, seeds, u_host,
                                                        mp, logp, logq, draws);
        case PFB_MODEL_FUNNEL:
            return launch_k3_m<KP, PFB_MODEL_FUNNEL>(st, n, K, nslots, unit_list, FR, HDR, seeds, u_host, mp,
                                                     logp, logq, draws);
        case PFB_MODEL_DIAGNORMAL:
            return launch_k3_m<KP, PFB_MODEL_DIAGNORMAL>(st, n, K, nslots, unit_list, FR, HDR, seeds, u_host,
                                                         mp, logp, logq, draws);
    }
    return cudaErrorInvalidValue;
}

// u_host != NULL: parity mode (normals [unit][K][n] supplied by the host).
// draws != NULL: materialise x (mode M / K5), [slot][K][n] i.e. column-major n x K per slot.
extern "C" cudaError_t PFB_K3_ENTRY(cudaStream_t st, int model, int n, int K, int nslots,
                                    const int32_t* unit_list, const double* FR, const double* HDR,
                                    const uint64_t* seeds, const double* u_host, const double* mp0,
                                    const double* mp1, double mc0, double* logp, double* logq,
                                    double* draws) {
    pfb_model_params mp{mp0, mp1, mc0};
    return launch_k3_k<PFB_K3_KP>(st, model, n, K, nslots, unit_list, FR, HDR, seeds, u_host, mp, logp, logq,
                                  draws);
}
