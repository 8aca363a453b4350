// K8 generic_logp — target log densities that need a GEMM over materialised draws (SURVEY §8d
// configs 4 and 5; the per-column closure call logp.(eachcol(draws)) of src/elbo.jl:15).
//
//   dense normal      logp(x) = -(x - m)' P (x - m) / 2        docs/src/examples/quickstart.md:17-24
//                     Y = P X by cuBLAS DGEMM, then per column  x'(Y - 2 P m) + m'P m
//   hier. logistic    theta = (log tau, b0, b_1..b_{n-2});  log tau ~ N(0,1), b0 ~ N(0, 2.5^2),
//                     b_j ~ N(0, tau^2), y_i ~ Bernoulli(sigmoid(b0 + x_i'b))          SURVEY §8d cfg 4
//                     eta = Xmat b by cuBLAS DGEMM, then per column the Bernoulli log likelihood
//
// The GEMMs are plain library GEMMs (cuBLAS); the kernels here are the fused epilogues: one warp
// per draw, lane-strided sums + xor butterfly (fixed order => deterministic).
#include "pfb_common.cuh"

#define PFB_K8_THREADS 256

// logp[col] = -( sum_i x_i (y_i - 2 pm_i) + mPm ) / 2;  slots with unit < 0 get NaN
__global__ void __launch_bounds__(PFB_K8_THREADS)
pfb_k8_dense_normal(int n, int64_t M, int K, const int32_t* __restrict__ slot_unit, const double* __restrict__ X,
                    const double* __restrict__ Y, const double* __restrict__ pm, double mPm,
                    double* __restrict__ logp) {
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * (PFB_K8_THREADS / 32) + (threadIdx.x >> 5);
    if (col >= M) return;
    if (slot_unit && slot_unit[col / K] < 0) {
        if (lane == 0) logp[col] = NAN;
        return;
    }
    const double* x = X + col * n;
    const double* y = Y + col * n;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s = fma(x[i], fma(-2.0, pm[i], y[i]), s);
    s = pfb_warp_sum(s);
    if (lane == 0) logp[col] = (s + mPm) / -2.0;
}

__device__ __forceinline__ double pfb_softplus(double e) {  // log(1 + exp(e)), stable
    return fmax(e, 0.0) + log1p(exp(-fabs(e)));
}

// ETA [nobs x M] = Xmat * beta (without intercept).  logp = prior(theta) + sum_i y_i eta_i - softplus(eta_i)
__global__ void __launch_bounds__(PFB_K8_THREADS)
pfb_k8_hier_logistic(int n, int nobs, int64_t M, int K, const int32_t* __restrict__ slot_unit,
                     const double* __restrict__ X, const double* __restrict__ ETA,
                     const double* __restrict__ yobs, double* __restrict__ logp) {
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * (PFB_K8_THREADS / 32) + (threadIdx.x >> 5);
    if (col >= M) return;
    if (slot_unit && slot_unit[col / K] < 0) {
        if (lane == 0) logp[col] = NAN;
        return;
    }
    const double* x = X + col * n;
    const double* eta = ETA + col * nobs;
    const double ltau = x[0], b0 = x[1];
    double ll = 0.0, bb = 0.0;
    for (int i = lane; i < nobs; i += 32) {
        const double e = eta[i] + b0;
        ll += fma(yobs[i], e, -pfb_softplus(e));
    }
    for (int j = 2 + lane; j < n; j += 32) bb = fma(x[j], x[j], bb);
    ll = pfb_warp_sum(ll);
    bb = pfb_warp_sum(bb);
    if (lane == 0) {
        const int nb = n - 2;
        const double half_log2pi = 0.5 * PFB_LOG2PI;
        double lp = -0.5 * ltau * ltau - half_log2pi;                                   // log tau ~ N(0, 1)
        lp += -0.5 * (b0 / 2.5) * (b0 / 2.5) - log(2.5) - half_log2pi;                  // b0 ~ N(0, 2.5^2)
        lp += -0.5 * bb * exp(-2.0 * ltau) - (double)nb * ltau - (double)nb * half_log2pi;  // b_j ~ N(0, tau^2)
        logp[col] = lp + ll;
    }
}

extern "C" cudaError_t pfb_launch_k8_dense(cudaStream_t st, int n, int64_t M, int K, const int32_t* slot_unit,
                                           const double* X, const double* Y, const double* pm, double mPm,
                                           double* logp) {
    if (M <= 0) return cudaSuccess;
    const int wpb = PFB_K8_THREADS / 32;
    pfb_k8_dense_normal<<<(unsigned)((M + wpb - 1) / wpb), PFB_K8_THREADS, 0, st>>>(n, M, K, slot_unit, X, Y, pm, mPm,
                                                                                    logp);
    return cudaGetLastError();
}

extern "C" cudaError_t pfb_launch_k8_logistic(cudaStream_t st, int n, int nobs, int64_t M, int K,
                                              const int32_t* slot_unit, const double* X, const double* ETA,
                                              const double* yobs, double* logp) {
    if (M <= 0) return cudaSuccess;
    const int wpb = PFB_K8_THREADS / 32;
    pfb_k8_hier_logistic<<<(unsigned)((M + wpb - 1) / wpb), PFB_K8_THREADS, 0, st>>>(n, nobs, M, K, slot_unit, X, ETA,
                                                                                     yobs, logp);
    return cudaGetLastError();
}
