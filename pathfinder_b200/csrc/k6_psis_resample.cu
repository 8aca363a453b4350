// K6 psis_pool + K7 resample_gather.
//
// K6 replaces PSIS.psis(log_ratios) as called from _compute_psis_result (reference:
// src/resample.jl:74-79).  PSIS.jl is not in the reference tree; this is the published
// algorithm (Vehtari et al., PSIS; Zhang & Stephens 2009 GPD fit with the weakly informative
// shape prior) exactly as restated in oracle/psis.py — the two are written operation for
// operation alike (same pf_math.h transcendentals, same canonical summation orders, this file
// is compiled with -fmad=false), so weights are bit-identical to the oracle's.
//
// K7 replaces _resample (src/resample.jl:58-72): inverse-CDF sampling on the fixed-point (2^52)
// cumulative weight table with 64 Philox bits per draw, column gather, component ids
// cld(ind, K_run).  Indices are 1-based like the reference's.
//
// K6 is latency-bound (N = pool size ~ 64 k): a single CTA of 1024 threads keeps every
// reduction in one deterministic order.  Top-(M+1) selection = 12-pass MSB radix select on the
// composite key (ordered double bits, index), then a shared-memory bitonic sort of the M+1
// candidates; M + 1 <= 8192 (N <= ~7.4 M).
#include "pfb_common.cuh"
#include "pf_rng.h"

#define PFB_K6_THREADS 1024
#define PFB_K6_MAXCAND 8192

__device__ __forceinline__ uint64_t pfb_ordered_key(double x) {
    uint64_t u = (uint64_t)__double_as_longlong(x);
    if (x != x) return 0xFFFFFFFFFFFFFFFFull;  // NaN sorts last (Julia isless)
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// canonical 32-lane sum: lane-strided sequential partials (done by the caller), xor-butterfly
__device__ __forceinline__ double pfb_butterfly32(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

struct pfb_psis_scalars {
    double pareto_k;
    double lse;
    double sigma;
    double logu;
    uint64_t Z;
    int64_t tail_len;
    int64_t smoothed;  // 1 if the tail was replaced
    uint64_t maxkey;   // ordered key of max(log weights) (device-internal)
};

// ---- K6 in stages.  Everything O(N) and elementwise runs grid-wide; everything whose floating-
// point ORDER is part of the contract with oracle/psis.py (the sums) keeps its single-CTA shape:
// 1024 lane-strided sequential partials, xor butterflies.  One CTA doing all of it was bound by a
// single SM's instruction rate (0.5 ms at N = 64 k, 3.3 ms at the 8-GPU pool N = 512 k).
#include <cub/device/device_radix_sort.cuh>

// K6a: log ratios, their order-preserving integer keys and indices (grid-wide)
__global__ void pfb_k6a_keys(int N, const double* __restrict__ logp, const double* __restrict__ logq,
                             const double* __restrict__ logr_in, double* __restrict__ logw,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double v = logr_in ? logr_in[i] : (logp[i] - logq[i]);
    // A NaN log ratio (a draw whose target or fitted density is undefined, e.g. from a failed path)
    // gets zero weight instead of poisoning the normalising sum and with it every weight.
    if (v != v) v = -INFINITY;
    logw[i] = v;
    keys[i] = pfb_ordered_key(v);
    idx[i] = (uint32_t)i;
}

// K6b: one CTA.  The last M + 1 entries of the ascending (key, index) sort are the cutoff element
// and the tail (ascending) — a stable radix sort orders ties by index, which is the composite order
// of the contract.  GPD fit (Zhang & Stephens) and tail replacement; writes the scalars.
__global__ void __launch_bounds__(PFB_K6_THREADS)
pfb_k6b_fit(int N, int M, int m_grid, const uint32_t* __restrict__ sorted_idx, double* __restrict__ logw,
            pfb_psis_scalars* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sX = reinterpret_cast<double*>(smem_raw);                   // MAXCAND   tail sample x
    uint32_t* cIdx = reinterpret_cast<uint32_t*>(sX + PFB_K6_MAXCAND);  // MAXCAND
    __shared__ double sB[128], sKj[128], sLL[128], sWj[128];
    __shared__ int sAllFinite;
    __shared__ double sScal[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = PFB_K6_THREADS;
    double pareto_k = NAN, sigma = NAN, logu = NAN;
    int smoothed = 0;
    if (M >= 5) {
        // cIdx[0] = cutoff element, cIdx[1..M] = tail ascending
        for (int t = tid; t < M + 1; t += NT) cIdx[t] = sorted_idx[N - (M + 1) + t];
        if (tid == 0) sAllFinite = 1;
        __syncthreads();
        for (int t = tid; t < M; t += NT) {
            const double v = logw[cIdx[1 + t]];
            if (!(fabs(v) <= 1.7976931348623157e308)) sAllFinite = 0;  // NaN or Inf
        }
        __syncthreads();
        logu = logw[cIdx[0]];
        if (sAllFinite) {
            const double lw_max = logw[cIdx[M]];
            const double mu_s = pf_exp(logu - lw_max);
            for (int t = tid; t < M; t += NT) sX[t] = pf_exp(logw[cIdx[1 + t]] - lw_max) - mu_s;
            __syncthreads();
            // ---- Zhang & Stephens empirical-Bayes fit ---------------------------------------------
            const double dM = (double)M;
            const double xmax = sX[M - 1];
            const double xq = sX[(int)floor(dM / 4.0 + 0.5) - 1];
            if (tid < m_grid) {
                const double jj = (double)(tid + 1);
                sB[tid] = 1.0 / xmax + (1.0 - sqrt((double)m_grid / (jj - 0.5))) / (3.0 * xq);
            }
            __syncthreads();
            for (int j = warp; j < m_grid; j += NT / 32) {
                const double b = sB[j];
                double acc = 0.0;
                for (int i = lane; i < M; i += 32) acc = acc + pf_log1p(-(b * sX[i]));
                acc = pfb_butterfly32(acc);
                if (lane == 0) {
                    const double kj = acc / dM;
                    sKj[j] = kj;
                    sLL[j] = dM * (pf_log(-(b / kj)) - kj - 1.0);
                }
            }
            __syncthreads();
            for (int j = warp; j < m_grid; j += NT / 32) {
                const double lj = sLL[j];
                double acc = 0.0;
                for (int i = lane; i < m_grid; i += 32) acc = acc + pf_exp(sLL[i] - lj);
                acc = pfb_butterfly32(acc);
                if (lane == 0) sWj[j] = 1.0 / acc;
            }
            __syncthreads();
            if (warp == 0) {
                double acc = 0.0;
                for (int i = lane; i < m_grid; i += 32) acc = acc + sWj[i];
                const double wsum = pfb_butterfly32(acc);
                acc = 0.0;
                for (int i = lane; i < m_grid; i += 32) acc = acc + sB[i] * (sWj[i] / wsum);
                const double b_post = pfb_butterfly32(acc);
                acc = 0.0;
                for (int i = lane; i < M; i += 32) acc = acc + pf_log1p(-(b_post * sX[i]));
                const double k_post = pfb_butterfly32(acc) / dM;
                if (lane == 0) {
                    sScal[0] = k_post;
                    sScal[1] = -k_post / b_post;                  // sigma
                    sScal[2] = (k_post * dM + 5.0) / (dM + 10.0);  // prior-adjusted shape
                }
            }
            __syncthreads();
            sigma = sScal[1];
            pareto_k = sScal[2];
            // ---- replace the tail by GPD quantiles ------------------------------------------------
            if (fabs(pareto_k) <= 1.7976931348623157e308) {
                smoothed = 1;
                for (int t = tid; t < M; t += NT) {
                    const double p = ((double)t + 0.5) / dM;
                    const double l1p = pf_log1p(-p);
                    double z;
                    if (fabs(pareto_k) < 2.220446049250313e-16) z = -l1p;
                    else z = pf_expm1(-(pareto_k * l1p)) / pareto_k;
                    const double qv = sigma * z;
                    const double val = pf_log(qv + mu_s);
                    const double mn = (val != val) ? val : (val < 0.0 ? val : 0.0);
                    logw[cIdx[1 + t]] = mn + lw_max;
                }
            }
        }
    }
    if (tid == 0) {
        out->pareto_k = pareto_k;
        out->sigma = sigma;
        out->logu = logu;
        out->tail_len = M;
        out->smoothed = smoothed;
        out->maxkey = pfb_ordered_key(-INFINITY);  // running maximum of K6c (NaN ignored)
    }
}

// K6c: maximum of the (smoothed) log weights, NaN ignored (grid-wide, exact in any order)
__global__ void pfb_k6c_max(int N, const double* __restrict__ logw, pfb_psis_scalars* __restrict__ out) {
    double mx = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        mx = fmax(mx, logw[i]);  // fmax ignores NaN
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0)
        atomicMax(reinterpret_cast<unsigned long long*>(&out->maxkey), (unsigned long long)pfb_ordered_key(mx));
}

__device__ __forceinline__ double pfb_key_to_double(uint64_t k) {
    const uint64_t u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)u);
}

// K6d: e_i = exp(logw_i - max) (grid-wide, elementwise)
__global__ void pfb_k6d_exp(int N, const double* __restrict__ logw, const pfb_psis_scalars* __restrict__ sc,
                            double* __restrict__ e) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    e[i] = pf_exp(logw[i] - pfb_key_to_double(sc->maxkey));
}

// K6e: one CTA — the canonical sum (1024 lane-strided sequential partials, butterflies) and lse
__global__ void __launch_bounds__(PFB_K6_THREADS)
pfb_k6e_sum(int N, const double* __restrict__ e, pfb_psis_scalars* __restrict__ out) {
    __shared__ double scratch[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc = 0.0;
    for (int i = tid; i < N; i += PFB_K6_THREADS) acc = acc + e[i];
    acc = pfb_butterfly32(acc);
    if (lane == 0) scratch[warp] = acc;
    __syncthreads();
    const double ssum = pfb_butterfly32(scratch[lane]);
    if (tid == 0) out->lse = pfb_key_to_double(out->maxkey) + pf_log(ssum);
}

// K6f: normalised log weights, weights, fixed-point (2^52) weights (grid-wide, elementwise)
__global__ void pfb_k6f_weights(int N, double* __restrict__ logw, double* __restrict__ weights,
                                uint64_t* __restrict__ cum, const pfb_psis_scalars* __restrict__ sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double lw = logw[i] - sc->lse;
    const double wv = pf_exp(lw);
    logw[i] = lw;
    weights[i] = wv;
    cum[i] = (wv > 0.0) ? (uint64_t)(wv * 4503599627370496.0) : 0ull;
}

// K6g: inclusive prefix sums of the integer weights (exact in any order), Z = total — grid-wide in three
// small launches (tile-local scans, scan of the tile totals, add the offsets).  The round-1 single-CTA
// version walked the table with a stride of N / 1024 per thread (uncoalesced): 0.62 ms of the 1.0 ms
// PSIS stage at the 8-GPU pool size (N = 512 k), 15 us now.
#define PFB_K6G_TILE 2048  // 256 threads x 8 consecutive elements
__global__ void __launch_bounds__(256)
pfb_k6g_tile_scan(int N, uint64_t* __restrict__ cum, uint64_t* __restrict__ tile_sums) {
    __shared__ uint64_t sW[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * PFB_K6G_TILE + (int64_t)tid * 8;
    uint64_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (base + k < N) ? cum[base + k] : 0ull;
#pragma unroll
    for (int k = 1; k < 8; ++k) v[k] += v[k - 1];
    uint64_t incl = v[7];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) sW[warp] = incl;
    __syncthreads();
    uint64_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += sW[w];
    const uint64_t excl = woff + incl - v[7];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (base + k < N) cum[base + k] = v[k] + excl;
    if (tid == 255) tile_sums[blockIdx.x] = excl + v[7];
}
// one CTA: exclusive scan of the tile totals in place (chunks of 1024 with a carry), Z = grand total
__global__ void __launch_bounds__(PFB_K6_THREADS)
pfb_k6g_sums_scan(int ntiles, uint64_t* __restrict__ tile_sums, pfb_psis_scalars* __restrict__ out) {
    __shared__ uint64_t sW[PFB_K6_THREADS / 32];
    __shared__ uint64_t sCarry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sCarry = 0ull;
    __syncthreads();
    for (int c0 = 0; c0 < ntiles; c0 += PFB_K6_THREADS) {
        const int i = c0 + tid;
        const uint64_t x = i < ntiles ? tile_sums[i] : 0ull;
        uint64_t incl = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) sW[warp] = incl;
        __syncthreads();
        uint64_t woff = sCarry;
        for (int w = 0; w < warp; ++w) woff += sW[w];
        if (i < ntiles) tile_sums[i] = woff + incl - x;
        __syncthreads();
        if (tid == PFB_K6_THREADS - 1) sCarry = woff + incl;
        __syncthreads();
    }
    if (tid == 0) out->Z = sCarry;
}
__global__ void pfb_k6g_add_offsets(int N, uint64_t* __restrict__ cum, const uint64_t* __restrict__ tile_offsets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) cum[i] += tile_offsets[i / PFB_K6G_TILE];
}

// K7: one CTA per output draw.  cum == nullptr => uniform sampling (importance = false).
__global__ void __launch_bounds__(128)
pfb_k7_resample_gather(int n, int N, int K_run, uint64_t seed, const uint64_t* __restrict__ cum,
                       const pfb_psis_scalars* __restrict__ sc, const double* __restrict__ pool,
                       int64_t* __restrict__ inds, int64_t* __restrict__ ids,
                       double* __restrict__ draws_out) {
    __shared__ int64_t sIdx;
    const int t = blockIdx.x;
    if (threadIdx.x == 0) {
        const uint64_t bits = pf_resample_bits((uint64_t)t, (uint32_t)seed, (uint32_t)(seed >> 32));
        int64_t idx;
        const uint64_t Z = (cum != nullptr) ? sc->Z : 0ull;
        if (Z == 0ull) {
            idx = (int64_t)pf_mulhi64(bits, (uint64_t)N);
        } else {
            const uint64_t target = pf_mulhi64(bits, Z);
            int lo = 0, hi = N - 1;  // first i with cum[i] > target
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cum[mid] > target) hi = mid; else lo = mid + 1;
            }
            idx = lo;
        }
        sIdx = idx;
        inds[t] = idx + 1;
        ids[t] = idx / K_run + 1;  // cld(idx + 1, K_run)
    }
    __syncthreads();
    if (pool != nullptr && draws_out != nullptr) {
        const double* src = pool + sIdx * (int64_t)n;
        double* dst = draws_out + (int64_t)t * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
}

extern "C" size_t pfb_psis_scalars_size() { return sizeof(pfb_psis_scalars); }

// Workspace of K6: keys in/out (2 N u64), indices in/out (2 N u32), e (N doubles), CUB scratch.
static size_t k6_cub_bytes(int N) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, N);
    return bytes;
}
extern "C" size_t pfb_k6_workspace_bytes(int N) {
    return (size_t)N * (16 + 8 + 8) + 256 + k6_cub_bytes(N);
}

extern "C" cudaError_t pfb_launch_k6(cudaStream_t st, int N, int M, int m_grid, const double* logp,
                                     const double* logq, const double* logr, double* logw, double* weights,
                                     uint64_t* cum, void* scalars, void* work, size_t work_bytes) {
    if (N <= 0) return cudaErrorInvalidValue;
    if (M + 1 > PFB_K6_MAXCAND || m_grid > 128) return cudaErrorInvalidValue;
    if (work_bytes < pfb_k6_workspace_bytes(N)) return cudaErrorInvalidValue;
    pfb_psis_scalars* sc = (pfb_psis_scalars*)scalars;
    uint64_t* k_in = (uint64_t*)work;
    uint64_t* k_out = k_in + N;
    double* e = (double*)(k_out + N);
    uint32_t* i_in = (uint32_t*)(e + N);
    uint32_t* i_out = i_in + N;
    void* tmp = (void*)(((uintptr_t)(i_out + N) + 255) & ~(uintptr_t)255);
    size_t tmp_bytes = k6_cub_bytes(N);
    const int TB = 256, GB = (N + TB - 1) / TB;
    pfb_k6a_keys<<<GB, TB, 0, st>>>(N, logp, logq, logr, logw, k_in, i_in);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    if (M >= 5) {
        err = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, i_in, i_out, N, 0, 64, st);
        if (err != cudaSuccess) return err;
    }
    const size_t smem = (size_t)PFB_K6_MAXCAND * (8 + 4);
    err = cudaFuncSetAttribute(pfb_k6b_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    pfb_k6b_fit<<<1, PFB_K6_THREADS, smem, st>>>(N, M, m_grid, i_out, logw, sc);
    const int GR = GB < 1184 ? GB : 1184;  // 148 SMs x 8
    pfb_k6c_max<<<GR, TB, 0, st>>>(N, logw, sc);
    pfb_k6d_exp<<<GB, TB, 0, st>>>(N, logw, sc, e);
    pfb_k6e_sum<<<1, PFB_K6_THREADS, 0, st>>>(N, e, sc);
    pfb_k6f_weights<<<GB, TB, 0, st>>>(N, logw, weights, cum, sc);
    {
        // e[] is free again: it hosts the tile totals of the prefix scan
        uint64_t* tile_sums = reinterpret_cast<uint64_t*>(e);
        const int ntiles = (N + PFB_K6G_TILE - 1) / PFB_K6G_TILE;
        pfb_k6g_tile_scan<<<ntiles, 256, 0, st>>>(N, cum, tile_sums);
        pfb_k6g_sums_scan<<<1, PFB_K6_THREADS, 0, st>>>(ntiles, tile_sums, sc);
        pfb_k6g_add_offsets<<<GB, TB, 0, st>>>(N, cum, tile_sums);
    }
    return cudaGetLastError();
}

extern "C" cudaError_t pfb_launch_k7(cudaStream_t st, int n, int N, int K_run, uint64_t seed, int ndraws,
                                     const uint64_t* cum, const void* scalars, const double* pool,
                                     int64_t* inds, int64_t* ids, double* draws_out) {
    if (ndraws <= 0) return cudaSuccess;
    pfb_k7_resample_gather<<<ndraws, 128, 0, st>>>(n, N, K_run, seed, cum, (const pfb_psis_scalars*)scalars,
                                                   pool, inds, ids, draws_out);
    return cudaGetLastError();
}

// ---- K7b: weighted sampling WITHOUT replacement (replace = false, src/resample.jl:58-72) -----------
// StatsBase.sample(rng, 1:N, pweights, ndraws; replace = false) is third-party (parity unpinned,
// it runs Efraimidis-Spirakis A-ExpJ on Julia's RNG stream).  The engine's contract is
// the same sampling design in its order-statistics form: every pool entry i gets the key
//     key_i = log(E_i) - log w_i,   E_i = -log(u_i) ~ Exp(1),  u_i from 53 Philox bits of counter i,
// and the ndraws smallest keys are the sample, in ascending key order (= the order sequential
// weighted sampling without replacement would have drawn them); ties break towards the smaller
// index.  importance = false: log w = 0 (a uniform random subset in random order).  Zero / NaN
// weights get key = +Inf and are taken only when nothing else is left.  pf_log makes the keys
// bit-identical to the oracle's (oracle/psis.py::resample_indices_norep).

__global__ void pfb_k7b_keys(int N, uint64_t seed, const double* __restrict__ logw, uint64_t* __restrict__ keys,
                             int32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint64_t bits = pf_resample_bits((uint64_t)i, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u = ((double)(bits >> 11) + 0.5) * 1.1102230246251565e-16;  // (0, 1), 2^-53 grid
    const double e = -pf_log(u);
    double lw = logw ? logw[i] : 0.0;
    double key = (lw == lw && lw > -INFINITY) ? pf_log(e) - lw : (double)INFINITY;
    keys[i] = pfb_ordered_key(key);
    idx[i] = i;
}

__global__ void __launch_bounds__(128)
pfb_k7b_gather(int n, int K_run, const int32_t* __restrict__ order, const double* __restrict__ pool,
               int64_t* __restrict__ inds, int64_t* __restrict__ ids, double* __restrict__ draws_out) {
    const int t = blockIdx.x;
    const int64_t idx = order[t];
    if (threadIdx.x == 0) {
        inds[t] = idx + 1;
        ids[t] = idx / K_run + 1;
    }
    if (pool != nullptr && draws_out != nullptr) {
        const double* src = pool + idx * (int64_t)n;
        double* dst = draws_out + (int64_t)t * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
}

// temp_bytes == NULL-query convention of CUB: call once with tmp = NULL to get the size.
extern "C" cudaError_t pfb_k7b_temp_bytes(int N, size_t* bytes) {
    *bytes = 0;
    return cub::DeviceRadixSort::SortPairs(nullptr, *bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                           (const int32_t*)nullptr, (int32_t*)nullptr, N);
}

// work: 2 N uint64 keys + 2 N int32 indices (in / out) laid out by the caller; tmp: CUB scratch.
extern "C" cudaError_t pfb_launch_k7b(cudaStream_t st, int n, int N, int K_run, uint64_t seed, int ndraws,
                                      const double* logw, const double* pool, uint64_t* keys_in, uint64_t* keys_out,
                                      int32_t* idx_in, int32_t* idx_out, void* tmp, size_t tmp_bytes, int64_t* inds,
                                      int64_t* ids, double* draws_out) {
    if (ndraws <= 0 || N <= 0) return cudaSuccess;
    pfb_k7b_keys<<<(N + 255) / 256, 256, 0, st>>>(N, seed, logw, keys_in, idx_in);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, idx_in, idx_out, N, 0, 64, st);
    if (e != cudaSuccess) return e;
    pfb_k7b_gather<<<ndraws, 128, 0, st>>>(n, K_run, idx_out, pool, inds, ids, draws_out);
    return cudaGetLastError();
}

// ---- K7r: bin the resampled pool indices by path, so that K3 can regenerate exactly those draws ----
// inds[t] (1-based pool index; the pool of this engine covers [base, base + P * K)) -> per path p the
// list of (draw k, output column t) in increasing t (one warp per path: ballot + prefix count, so the
// lists are deterministic).  Entries of other ranks' paths are skipped.
__global__ void pfb_k7r_bin(int P, int K, int m, int64_t base, const int64_t* __restrict__ inds,
                            int32_t* __restrict__ cnt, int2* __restrict__ list) {
    const int p = blockIdx.x, lane = threadIdx.x;
    int c = 0;
    for (int t0 = 0; t0 < m; t0 += 32) {
        const int t = t0 + lane;
        int64_t loc = -1;
        if (t < m) loc = inds[t] - 1 - base;
        const bool mine = (loc >= 0) && (loc / K == p) && (loc < (int64_t)P * K);
        const unsigned bal = __ballot_sync(0xffffffffu, mine);
        if (mine) list[(int64_t)p * m + c + __popc(bal & ((1u << lane) - 1u))] = make_int2((int)(loc % K), t);
        c += __popc(bal);
    }
    if (lane == 0) cnt[p] = c;
}

extern "C" cudaError_t pfb_launch_k7r_bin(cudaStream_t st, int P, int K, int m, int64_t base, const int64_t* inds,
                                          int32_t* cnt, void* list) {
    if (P <= 0 || m <= 0) return cudaSuccess;
    pfb_k7r_bin<<<P, 32, 0, st>>>(P, K, m, base, inds, cnt, reinterpret_cast<int2*>(list));
    return cudaGetLastError();
}
