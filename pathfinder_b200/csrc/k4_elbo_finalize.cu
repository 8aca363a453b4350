// K4 elbo_finalize_argmax — K4a: one warp per unit (grid over units); K4b: one thread per path.
//
// Replaces the reduction part of elbo_and_samples (reference: src/elbo.jl:16-18):
//   logr = logp - logq;  elbo = mean(logr);  se = sqrt(var(logr; mean = elbo) / K)   (var: K-1)
// then the NaN-skipping argmax of maximize_elbo (src/elbo.jl:7-9, src/utils.jl:57-72) and the
// per-path success flag (src/singlepath.jl:297-314).
// One warp per unit (lane-strided partial sums + xor-butterfly, fixed order => deterministic);
// two passes over the K values so the variance is the reference's two-pass formula.
#include "pfb_common.cuh"

#define PFB_K4_THREADS 256

// K4a: one warp per unit, 8 units per CTA (grid over all units: the whole GPU takes part)
__global__ void __launch_bounds__(PFB_K4_THREADS)
pfb_k4a_unit_stats(int K, int64_t U, const double* __restrict__ logp, const double* __restrict__ logq,
                   double* __restrict__ elbo, double* __restrict__ se) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t u = (int64_t)blockIdx.x * nw + warp;
    if (u >= U) return;
    const double* lp = logp + u * (int64_t)K;
    const double* lq = logq + u * (int64_t)K;
    double s = 0.0;
    for (int k = lane; k < K; k += 32) s += lp[k] - lq[k];
    s = pfb_warp_sum(s);
    const double mean = s / (double)K;
    double v = 0.0;
    for (int k = lane; k < K; k += 32) {
        double d = (lp[k] - lq[k]) - mean;
        v = fma(d, d, v);
    }
    v = pfb_warp_sum(v);
    if (lane == 0) {
        elbo[u] = mean;
        se[u] = sqrt(v / (double)(K - 1) / (double)K);
    }
}

// K4b: one thread per path — the NaN-skipping argmax and the success flag
__global__ void pfb_k4b_argmax(int P, const int64_t* __restrict__ point_off, const double* __restrict__ elbo,
                               int64_t* __restrict__ best_iter, int32_t* __restrict__ best_unit,
                               int32_t* __restrict__ success) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int64_t c0 = point_off[p];
    const int L = (int)(point_off[p + 1] - c0) - 1;
    const int64_t u0 = c0 - p;
    // _findmax_skipnan: the first value is kept even if NaN until a non-NaN arrives; later
    // NaNs are skipped; ties keep the earliest.
    int best = 0;  // 1-based, 0 = empty
    double bv = NAN;
    for (int l = 0; l < L; ++l) {
        double x = elbo[u0 + l];
        if (l == 0) { best = 1; bv = x; continue; }
        if (x != x) continue;
        if (bv != bv || x > bv) { bv = x; best = l + 1; }
    }
    best_iter[p] = best;
    const bool ok = (L > 0) && !(bv != bv) && (bv != -INFINITY);
    success[p] = ok ? 1 : 0;
    best_unit[p] = (best > 0) ? (int32_t)(u0 + best - 1) : -1;
}

extern "C" cudaError_t pfb_launch_k4(cudaStream_t st, int P, int K, int64_t U, const int64_t* point_off,
                                     const double* logp, const double* logq, double* elbo, double* se,
                                     int64_t* best_iter, int32_t* best_unit, int32_t* success) {
    if (P <= 0) return cudaSuccess;
    if (U > 0) {
        const int per = PFB_K4_THREADS / 32;
        pfb_k4a_unit_stats<<<(unsigned)((U + per - 1) / per), PFB_K4_THREADS, 0, st>>>(K, U, logp, logq, elbo, se);
    }
    pfb_k4b_argmax<<<(P + 127) / 128, 128, 0, st>>>(P, point_off, elbo, best_iter, best_unit, success);
    return cudaGetLastError();
}
