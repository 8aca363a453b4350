// K8g gemm_logp — the GEMM-shaped target log densities (BASELINE configs 4 and 5) as ONE kernel: an FP64
// tensor-core GEMM over the materialised draws whose epilogue reduces straight to log p, so neither
// the product matrix (n x M doubles — 16 MB per unit at config 5) nor a library GEMM is involved.
//
//   dense normal    logp(x) = -(x - m)' P (x - m) / 2 = -( x'(P x - 2 P m) + m'P m ) / 2
//                   docs/src/examples/quickstart.md:17-24; the closure call logp.(eachcol(draws)), src/elbo.jl:15
//   hier. logistic  theta = (log tau, b0, b_1..b_{n-2}); eta = Xmat b; log p = prior + sum_i y_i e_i - softplus(e_i),
//                   e_i = eta_i + b0                                                         SURVEY §8d config 4
//
// C[Mr x 128 draws] = A[Mr x Kd] * B[Kd x 128 draws] per CTA, row block by row block (128 rows), on
// mma.sync.m8n8k4.f64 (SASS DMMA — FP64 has no tcgen05 kind): 16 warps, warp tile 32 x 32 (16 DMMA per
// 8 fragment loads; four warps per scheduler keep the FP64 pipe fed), 32-deep k-tiles double-buffered through
// cp.async in shared memory with padded leading dimensions (132 / 36 doubles) that make both fragment
// patterns bank-conflict free (16-deep tiles in a 3-stage ring spend twice as long at barriers).
// After the last k-tile of a row block the accumulators are folded into per-draw sums
// (x .* (y - 2 P m), or the Bernoulli terms) and cleared; a CTA owns its 128 draws for ALL rows, so
// every log p is written once, in a fixed summation order (deterministic).
#include "pfb_common.cuh"

#define K8G_BM 128
#define K8G_BN 128
#ifndef K8G_BK
#define K8G_BK 32
#endif
#define K8G_LDA (K8G_BM + 4)
#define K8G_LDB (K8G_BK + 4)
#ifndef K8G_STAGES
#define K8G_STAGES 2
#endif
#ifndef K8G_WARPS_M
#define K8G_WARPS_M 4  // warps along the rows of the 128 x 128 tile (x 4 along the draws): 16 warps, warp tile 32 x 32
#endif
#define K8G_MI (K8G_BM / K8G_WARPS_M / 8)  // 8 x 8 row tiles per warp
#define K8G_THREADS (K8G_WARPS_M * 4 * 32)
#define K8G_CPT (K8G_BK * 64 / K8G_THREADS)  // 16-byte chunks per thread, operand and tile
#define K8G_BCH (K8G_BK / 2)                 // 16-byte chunks per draw of a B tile
static_assert((K8G_LDA % 16) == 4 && (K8G_LDB % 16) == 4, "padded leading dimensions: conflict-free fragment loads");
static_assert(K8G_STAGES >= 2, "at least double buffering");
#define K8G_STAGE_DOUBLES (K8G_BK * K8G_LDA + K8G_BN * K8G_LDB)

__device__ __forceinline__ void k8g_cp_async8(double* smem_dst, const double* gmem_src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;  // src-size 0: zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void k8g_cp_async16(uint32_t smem_dst, const double* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void k8g_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void k8g_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ double k8g_softplus(double e) {  // log(1 + exp(e)), stable
    return fmax(e, 0.0) + log1p(exp(-fabs(e)));
}

struct k8g_params {
    const double* A;    // Mr x Kd, column-major, leading dimension lda (P, or the design matrix)
    int lda;
    const double* X;    // draws, n x Ncols column-major (one draw per column)
    int n;
    int brow0;          // first draw row that enters the product (0: dense normal; 2: coefficients of the logistic model)
    int Mr, Kd;
    int64_t Ncols;
    int K;              // draws per slot (for slot_unit)
    const int32_t* slot_unit;  // nullptr, or slots with unit < 0 get NaN
    const double* v0;   // dense: P m [n];  logistic: y [nobs]
    double c0;          // dense: m' P m
    double* logp;
};

// MODEL 0: dense normal, 1: hierarchical logistic regression
template <int MODEL>
__global__ void __launch_bounds__(K8G_THREADS, 1) pfb_k8_gemm_logp(k8g_params p) {
    extern __shared__ __align__(16) double k8g_smem[];
    __shared__ double sPart[K8G_WARPS_M][K8G_BN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % K8G_WARPS_M, wn = warp / K8G_WARPS_M;
    const int64_t d0 = (int64_t)blockIdx.x * K8G_BN;
    const int nkt = (p.Kd + K8G_BK - 1) / K8G_BK;
    const int nrb = (p.Mr + K8G_BM - 1) / K8G_BM;
    const int ntiles = nkt * nrb;

    // Interior tiles of 16-byte-aligned operands take the cheap path: four 16-byte cp.async per operand and
    // thread from per-thread base pointers (the generic path below spends ~19 integer instructions per
    // 8-byte copy on addresses and predicates, which cost 28 % of the tensor pipe's time).
    const bool aligned = ((p.lda & 1) == 0) && ((p.n & 1) == 0) && ((p.brow0 & 1) == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0);
    const bool cols_full = d0 + K8G_BN <= p.Ncols;
    // A chunk (k = tid / 64 + (T / 64) i, m = 2 (tid % 64)), B chunk (d = tid / BCH + (T / BCH) i, k = 2 (tid % BCH))
    const double* a_thr = p.A + (int64_t)(tid >> 6) * p.lda + 2 * (tid & 63);
    const double* b_thr = p.X + (d0 + (tid / K8G_BCH)) * (int64_t)p.n + p.brow0 + 2 * (tid % K8G_BCH);
    const uint32_t sa_thr = (uint32_t)__cvta_generic_to_shared(k8g_smem) + 8u * ((tid >> 6) * K8G_LDA + 2 * (tid & 63));
    const uint32_t sb_thr = (uint32_t)__cvta_generic_to_shared(k8g_smem) +
                            8u * (K8G_BK * K8G_LDA + (tid / K8G_BCH) * K8G_LDB + 2 * (tid % K8G_BCH));
    const int64_t a_step = (K8G_THREADS / 64) * (int64_t)p.lda, b_step = (K8G_THREADS / K8G_BCH) * (int64_t)p.n;

    auto load_tile = [&](int tile, int stage) {
        const int rb = tile / nkt, kt = tile - rb * nkt;
        const int m0 = rb * K8G_BM, k0 = kt * K8G_BK;
        if (aligned && cols_full && k0 + K8G_BK <= p.Kd && m0 + K8G_BM <= p.Mr) {
            const double* pa = a_thr + (int64_t)k0 * p.lda + m0;
            const double* pb = b_thr + k0;
            const uint32_t so = (uint32_t)stage * (K8G_STAGE_DOUBLES * 8u);
#pragma unroll
            for (int i = 0; i < K8G_CPT; ++i) {
                k8g_cp_async16(sa_thr + so + 8u * ((K8G_THREADS / 64) * i * K8G_LDA), pa + i * a_step);
                k8g_cp_async16(sb_thr + so + 8u * ((K8G_THREADS / K8G_BCH) * i * K8G_LDB), pb + i * b_step);
            }
            return;
        }
        double* sA = k8g_smem + (size_t)stage * K8G_STAGE_DOUBLES;
        double* sB = sA + K8G_BK * K8G_LDA;
#pragma unroll
        for (int i = 0; i < (K8G_BK * K8G_BM) / K8G_THREADS; ++i) {
            const int e = tid + K8G_THREADS * i;
            const int k = e / K8G_BM, m = e - k * K8G_BM;
            const bool ok = (k0 + k < p.Kd) && (m0 + m < p.Mr);
            const double* src = ok ? p.A + (int64_t)(k0 + k) * p.lda + (m0 + m) : p.A;
            k8g_cp_async8(sA + k * K8G_LDA + m, src, ok);
        }
#pragma unroll
        for (int i = 0; i < (K8G_BK * K8G_BN) / K8G_THREADS; ++i) {
            const int e = tid + K8G_THREADS * i;
            const int d = e / K8G_BK, k = e - d * K8G_BK;
            const bool ok = (k0 + k < p.Kd) && (d0 + d < p.Ncols);
            const double* src = ok ? p.X + (d0 + d) * (int64_t)p.n + p.brow0 + (k0 + k) : p.X;
            k8g_cp_async8(sB + d * K8G_LDB + k, src, ok);
        }
    };

    double acc[K8G_MI][4][2];
#pragma unroll
    for (int i = 0; i < K8G_MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double part[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) part[j][0] = part[j][1] = 0.0;

    // per-draw constants of the epilogue (this lane's 8 draws: d0 + wn 32 + j 8 + 2 t + c)
    double b0v[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int64_t d = d0 + wn * 32 + j * 8 + 2 * t + c;
            b0v[j][c] = (MODEL == 1 && d < p.Ncols) ? p.X[d * (int64_t)p.n + 1] : 0.0;
        }

#pragma unroll
    for (int s = 0; s < K8G_STAGES - 1; ++s) {
        if (s < ntiles) load_tile(s, s);
        k8g_commit();
    }
    for (int tile = 0; tile < ntiles; ++tile) {
        k8g_wait<K8G_STAGES - 2>();
        __syncthreads();  // tile `tile` has landed; every warp is done with the stage about to be refilled
        {
            const int nt = tile + K8G_STAGES - 1;
            if (nt < ntiles) load_tile(nt, nt % K8G_STAGES);
            k8g_commit();
        }
        const double* sA = k8g_smem + (size_t)(tile % K8G_STAGES) * K8G_STAGE_DOUBLES;
        const double* sB = sA + K8G_BK * K8G_LDA;
#pragma unroll
        for (int kk = 0; kk < K8G_BK; kk += 4) {
            double a[K8G_MI], b[4];
#pragma unroll
            for (int i = 0; i < K8G_MI; ++i) a[i] = sA[(kk + t) * K8G_LDA + wm * (K8G_MI * 8) + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[(wn * 32 + j * 8 + g) * K8G_LDB + kk + t];
#pragma unroll
            for (int i = 0; i < K8G_MI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                        : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                        : "d"(a[i]), "d"(b[j]));
        }
        const int rb = tile / nkt, kt = tile - rb * nkt;
        if (kt == nkt - 1) {
            // ---- the row block is complete: fold it into the per-draw sums ----------------------------
            const int m0 = rb * K8G_BM;
#pragma unroll
            for (int i = 0; i < K8G_MI; ++i) {
                const int m = m0 + wm * (K8G_MI * 8) + i * 8 + g;
                const bool mok = m < p.Mr;
                const double vm = mok ? p.v0[m] : 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int64_t d = d0 + wn * 32 + j * 8 + 2 * t + c;
                        if (mok && d < p.Ncols) {
                            if (MODEL == 0) {
                                const double x = p.X[d * (int64_t)p.n + m];
                                part[j][c] = fma(x, fma(-2.0, vm, acc[i][j][c]), part[j][c]);
                            } else {
                                const double e = acc[i][j][c] + b0v[j][c];
                                part[j][c] += fma(vm, e, -k8g_softplus(e));
                            }
                        }
                        acc[i][j][c] = 0.0;
                    }
            }
        }
    }
    k8g_wait<0>();
    // ---- reduce over the 8 row lanes (g) of the warp, then over the row warps ----------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double v = part[j][c];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) sPart[wm][wn * 32 + j * 8 + 2 * t + c] = v;
        }
    __syncthreads();
    if (tid < K8G_BN) {
        const int64_t d = d0 + tid;
        if (d < p.Ncols) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < K8G_WARPS_M; ++w) s += sPart[w][tid];
            double out;
            if (p.slot_unit && p.slot_unit[d / p.K] < 0) {
                out = NAN;
            } else if (MODEL == 0) {
                out = (s + p.c0) / -2.0;
            } else {
                const double* x = p.X + d * (int64_t)p.n;
                const double ltau = x[0], b0 = x[1];
                double bb = 0.0;
                for (int jx = 2; jx < p.n; ++jx) bb = fma(x[jx], x[jx], bb);
                const int nb = p.n - 2;
                const double half_log2pi = 0.5 * PFB_LOG2PI;
                double lp = -0.5 * ltau * ltau - half_log2pi;                                      // log tau ~ N(0, 1)
                lp += -0.5 * (b0 / 2.5) * (b0 / 2.5) - log(2.5) - half_log2pi;                     // b0 ~ N(0, 2.5^2)
                lp += -0.5 * bb * exp(-2.0 * ltau) - (double)nb * ltau - (double)nb * half_log2pi;  // b_j ~ N(0, tau^2)
                out = lp + s;
            }
            p.logp[d] = out;
        }
    }
}

template <int MODEL>
static cudaError_t k8g_launch(cudaStream_t st, const k8g_params& p) {
    if (p.Ncols <= 0) return cudaSuccess;
    const size_t smem = (size_t)K8G_STAGES * K8G_STAGE_DOUBLES * 8;
    cudaError_t e = cudaFuncSetAttribute(pfb_k8_gemm_logp<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t grid = (p.Ncols + K8G_BN - 1) / K8G_BN;
    if (grid > 2147483647LL) return cudaErrorInvalidValue;
    pfb_k8_gemm_logp<MODEL><<<(unsigned)grid, K8G_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

// X: draws [n x M] column-major.  Pd: precision matrix [n x n] column-major, pm = P m, mPm = m'P m.
extern "C" cudaError_t pfb_launch_k8g_dense(cudaStream_t st, int n, int64_t M, int K, const int32_t* slot_unit,
                                            const double* X, const double* Pd, const double* pm, double mPm,
                                            double* logp) {
    k8g_params p{Pd, n, X, n, 0, n, n, M, K, slot_unit, pm, mPm, logp};
    return k8g_launch<0>(st, p);
}
// Xmat: design matrix [nobs x (n - 2)] column-major, yobs [nobs].
extern "C" cudaError_t pfb_launch_k8g_logistic(cudaStream_t st, int n, int nobs, int64_t M, int K,
                                               const int32_t* slot_unit, const double* X, const double* Xmat,
                                               const double* yobs, double* logp) {
    k8g_params p{Xmat, nobs, X, n, 2, nobs, n - 2, M, K, slot_unit, yobs, 0.0, logp};
    return k8g_launch<1>(st, p);
}
