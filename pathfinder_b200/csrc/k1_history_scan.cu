// K1 lbfgs_history_scan — per path, sequential over the trajectory.
//
// Replaces the state machine of lbfgs_inverse_hessians (reference: src/inverse_hessian.jl:25-66):
// for l = 1..L: s = theta_l - theta_{l-1}, y = g_{l-1} - g_l (g = gradient of log density),
// curvature test y's > eps * y'y (:47), ring-buffer index mod1(ind+1, J) (:49),
// J_eff = max(ind, J_eff) (:50), Gilbert-Lemarechal diagonal update (:5-10, :55) or a rejected
// update (:57).  The inverse Hessian for iteration l is built *even when the update was
// rejected* (:61), so every unit gets (alpha, history) written.
//
// One CTA per path; threads own rows i = tid, tid + blockDim, ...; alpha lives in registers
// (<= PFB_K1_RPT rows per thread) or in global memory for very large n.  One block reduction of
// 4 dot products per iteration.  HBM-bound on reading 2 new columns (theta_l, g_l) per step and
// writing alpha: 8n*3 bytes per unit.
#include "pfb_common.cuh"

#define PFB_K1_RPT_WIDE 8   // up to 512 threads x 8 rows (n <= 4096): short links of the sequential chain
#define PFB_K1_RPT_TALL 16  // 256 threads x 16 rows otherwise; beyond that alpha lives in global memory

template <bool IN_REGS, int PFB_K1_RPT>
__global__ void __launch_bounds__(PFB_K1_RPT == PFB_K1_RPT_WIDE ? 512 : 256)
pfb_k1_history_scan(int n, int J, double eps, const double* __restrict__ X,
                    const double* __restrict__ G, const int64_t* __restrict__ point_off,
                    double* __restrict__ alpha_out,   // [n x U]
                    int32_t* __restrict__ hist,       // [U x J] point columns (global), oldest first
                    int32_t* __restrict__ hist_cnt,   // [U]
                    int64_t* __restrict__ n_rejected, // [P]
                    int p_base                        // first path of this launch (pipelined uploads)
) {
    __shared__ double scratch[4 * 32];
    __shared__ int32_t ring[64];  // J <= 20
    const int p = blockIdx.x + p_base;
    const int64_t c0 = point_off[p];
    const int L = (int)(point_off[p + 1] - c0) - 1;
    const int64_t u0 = c0 - p;  // first unit of this path
    const int tid = threadIdx.x, nt = blockDim.x;

    double a_reg[PFB_K1_RPT];
    if (IN_REGS) {
#pragma unroll
        for (int r = 0; r < PFB_K1_RPT; ++r) a_reg[r] = 1.0;
    }
    int ind = 0, jeff = 0;
    long long rejected = 0;

    for (int l = 1; l <= L; ++l) {
        const double* x0 = X + (c0 + l - 1) * (int64_t)n;
        const double* x1 = X + (c0 + l) * (int64_t)n;
        const double* g0 = G + (c0 + l - 1) * (int64_t)n;
        const double* g1 = G + (c0 + l) * (int64_t)n;
        double* a_out = alpha_out + (u0 + l - 1) * (int64_t)n;
        const double* a_prev = (l > 1) ? alpha_out + (u0 + l - 2) * (int64_t)n : nullptr;

        double v[4] = {0.0, 0.0, 0.0, 0.0};  // y's, y'y, y' diag(alpha) y, s' diag(alpha)^-1 s
        if (IN_REGS) {
#pragma unroll
            for (int r = 0; r < PFB_K1_RPT; ++r) {
                int i = tid + r * nt;
                if (i < n) {
                    double s = x1[i] - x0[i], y = g0[i] - g1[i], a = a_reg[r];
                    v[0] = fma(y, s, v[0]);
                    v[1] = fma(y, y, v[1]);
                    v[2] = fma(y * a, y, v[2]);
                    v[3] = fma(s / a, s, v[3]);
                }
            }
        } else {
            for (int i = tid; i < n; i += nt) {
                double s = x1[i] - x0[i], y = g0[i] - g1[i], a = a_prev ? a_prev[i] : 1.0;
                v[0] = fma(y, s, v[0]);
                v[1] = fma(y, y, v[1]);
                v[2] = fma(y * a, y, v[2]);
                v[3] = fma(s / a, s, v[3]);
            }
        }
        pfb_block_sum<4>(v, scratch);
        const bool accept = v[0] > eps * v[1];  // :47 (false for NaN)
        if (accept) {
            ind = ind % J + 1;          // mod1(ind + 1, J)
            jeff = max(ind, jeff);
            if (tid == 0) ring[ind - 1] = (int32_t)(c0 + l - 1);  // pair (point l-1 -> l)
            const double aa = v[2], b = v[0], c = v[3];
            const double aoc = aa / c;
            // alpha' = b / (a/alpha + y^2 - (a/c) (s/alpha)^2)        (:9)
            if (IN_REGS) {
#pragma unroll
                for (int r = 0; r < PFB_K1_RPT; ++r) {
                    int i = tid + r * nt;
                    if (i < n) {
                        double s = x1[i] - x0[i], y = g0[i] - g1[i], a = a_reg[r];
                        double sa = s / a;
                        a_reg[r] = b / (aa / a + y * y - aoc * (sa * sa));
                        a_out[i] = a_reg[r];
                    }
                }
            } else {
                for (int i = tid; i < n; i += nt) {
                    double s = x1[i] - x0[i], y = g0[i] - g1[i], a = a_prev ? a_prev[i] : 1.0;
                    double sa = s / a;
                    a_out[i] = b / (aa / a + y * y - aoc * (sa * sa));
                }
            }
        } else {
            rejected++;
            if (IN_REGS) {
#pragma unroll
                for (int r = 0; r < PFB_K1_RPT; ++r) {
                    int i = tid + r * nt;
                    if (i < n) a_out[i] = a_reg[r];
                }
            } else {
                for (int i = tid; i < n; i += nt) a_out[i] = a_prev ? a_prev[i] : 1.0;
            }
        }
        __syncthreads();  // ring[] visible; (global alpha of this step visible to next step)
        if (tid < jeff) {
            // hist_inds = [(ind+1):J_eff ; 1:ind]  (1-based ring slots, oldest first)   (:105)
            int slot = (tid < jeff - ind) ? (ind + tid) : (tid - (jeff - ind));
            hist[(u0 + l - 1) * (int64_t)J + tid] = ring[slot];
        }
        if (tid == 0) hist_cnt[u0 + l - 1] = jeff;
    }
    if (tid == 0) n_rejected[p] = rejected;
}

// paths [p_base, p_base + P)
extern "C" cudaError_t pfb_launch_k1_range(cudaStream_t st, int n, int p_base, int P, int J, double eps, const double* X,
                                           const double* G, const int64_t* point_off, double* alpha,
                                           int32_t* hist, int32_t* hist_cnt, int64_t* n_rejected);
extern "C" cudaError_t pfb_launch_k1(cudaStream_t st, int n, int P, int J, double eps, const double* X,
                                     const double* G, const int64_t* point_off, double* alpha,
                                     int32_t* hist, int32_t* hist_cnt, int64_t* n_rejected) {
    return pfb_launch_k1_range(st, n, 0, P, J, eps, X, G, point_off, alpha, hist, hist_cnt, n_rejected);
}
extern "C" cudaError_t pfb_launch_k1_range(cudaStream_t st, int n, int p_base, int P, int J, double eps, const double* X,
                                           const double* G, const int64_t* point_off, double* alpha,
                                           int32_t* hist, int32_t* hist_cnt, int64_t* n_rejected) {
    if (P <= 0) return cudaSuccess;
    // only P CTAs exist and each is a sequential chain over the trajectory: more threads per CTA
    // shorten every link (2 rows per thread at n = 1024)
    if (n <= 512 * PFB_K1_RPT_WIDE) {
        const int threads = n >= 512 ? 512 : 256;
        pfb_k1_history_scan<true, PFB_K1_RPT_WIDE><<<P, threads, 0, st>>>(n, J, eps, X, G, point_off, alpha, hist,
                                                                          hist_cnt, n_rejected, p_base);
    } else if (n <= 256 * PFB_K1_RPT_TALL) {
        pfb_k1_history_scan<true, PFB_K1_RPT_TALL><<<P, 256, 0, st>>>(n, J, eps, X, G, point_off, alpha, hist,
                                                                      hist_cnt, n_rejected, p_base);
    } else {
        pfb_k1_history_scan<false, PFB_K1_RPT_TALL><<<P, 256, 0, st>>>(n, J, eps, X, G, point_off, alpha, hist,
                                                                       hist_cnt, n_rejected, p_base);
    }
    return cudaGetLastError();
}
