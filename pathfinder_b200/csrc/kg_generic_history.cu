// Kg — the generic (runtime-width) pair K2g / K3g for history_length > 12.
//
// The reference accepts any history length (src/inverse_hessian.jl:25); the tensor-core kernels K2 / K3
// are instantiated for 2J <= 24 reflector columns (their record layouts, register-resident accumulators
// and shared-memory panels are sized at compile time).  Wider histories are rare and far from the
// bench's hot path, so they get one straightforward pair of kernels with the SAME algorithm and
// conventions (LAPACK dlarfg / dlarft Householder QR so that Q equals the reference's, dpotrf
// semantics for the Cholesky factor, the RNG contract of pf_rng.h) and run-time widths: every small
// matrix lives in a per-CTA global workspace, every long sum is a warp-strided dot product.
//
//   K2g  lbfgs_inverse_hessian + pdfactorize + logabsdet + mu      (src/inverse_hessian.jl:98-133,
//        src/woodbury.jl:201-207, :77-80, src/mvnormal.jl:17)
//   K3g  rand_and_logpdf + log p of the diagonal-quadratic families (src/mvnormal.jl:24-39, src/elbo.jl:15)
//
// Layouts: FRg [U][n][KP + 2] rows { Vh[i][0..KP), sqrt(alpha_i), mu_i } (the FR layout of
// pfb_common.cuh), HDR as for K2 (T and Vc row-major KP x KP, logdet, PD flag, k_eff).
#include "pfb_common.cuh"
#include "pf_rng.h"

#define KG_THREADS 256
#define KG_WARPS (KG_THREADS / 32)

// workspace doubles per CTA of K2g
__host__ __device__ inline size_t kg_ws_doubles(int KP) {
    const size_t k2 = (size_t)KP * KP, j2 = (size_t)(KP / 2) * (KP / 2);
    return 5 * k2 + 4 * j2 + 4 * (size_t)KP + 8;
}

__device__ __forceinline__ double kg_warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__global__ void __launch_bounds__(KG_THREADS)
pfb_k2g_woodbury_build(int KP, int n, int U, int u_base, int J, const double* __restrict__ X,
                       const double* __restrict__ G, const int32_t* __restrict__ unit_col,
                       const double* __restrict__ alpha_all, const int32_t* __restrict__ hist,
                       const int32_t* __restrict__ hist_cnt, double* __restrict__ FRg, double* __restrict__ HDR,
                       double* __restrict__ WS) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = KG_THREADS;
    const int RS = KP + 2, JM = KP / 2;
    double* ws = WS + (size_t)blockIdx.x * kg_ws_doubles(KP);
    double* wD = ws;                 // KP x KP
    double* wRq = wD + KP * KP;      // KP x KP
    double* wE = wRq + KP * KP;      // KP x KP
    double* wC = wE + KP * KP;       // KP x KP
    double* wVtV = wC + KP * KP;     // KP x KP
    double* wStY = wVtV + KP * KP;   // JM x JM
    double* wYaY = wStY + JM * JM;
    double* wNR = wYaY + JM * JM;
    double* wM = wNR + JM * JM;
    double* wDots = wM + JM * JM;    // KP
    double* wTau = wDots + KP;       // KP
    double* wW = wTau + KP;          // KP
    double* wW2 = wW + KP;           // KP
    __shared__ double sB[4];
    __shared__ int sFlag;

    for (int uu = blockIdx.x; uu < U; uu += gridDim.x) {
        const int u = uu + u_base;
        const int jeff = hist_cnt[u];
        const int kc = 2 * jeff, kq = min(n, kc);
        double* fr = FRg + (int64_t)u * n * RS;
        double* hdr = HDR + (int64_t)u * pfb_hs_of(KP);
        double* hT = hdr;            // T,  KP x KP row-major
        double* hVc = hdr + KP * KP; // Vc, KP x KP row-major
        const double* alpha = alpha_all + (int64_t)u * n;
        const int64_t col = unit_col[u];
        const double* theta = X + col * n;
        const double* g = G + col * n;
        const int32_t* hu = hist + (int64_t)u * J;
#define PNL(i, c) fr[(int64_t)(i) * RS + (c)]
        // ---- panel A~ = [ sqrt(alpha) .* Y | S ./ sqrt(alpha) ], extra columns sqrt(alpha), t = sqrt(alpha) g ----
        for (int i = tid; i < n; i += nt) {
            const double a = alpha[i], sa = sqrt(a);
            for (int j = 0; j < jeff; ++j) {
                const int64_t c = hu[j];
                const double s = X[(c + 1) * n + i] - X[c * n + i];
                const double y = G[c * n + i] - G[(c + 1) * n + i];
                PNL(i, j) = (a * y) / sa;
                PNL(i, jeff + j) = s / sa;
            }
            for (int j = kc; j < KP; ++j) PNL(i, j) = 0.0;
            PNL(i, KP) = sa;
            PNL(i, KP + 1) = sa * g[i];
        }
        for (int e = tid; e < KP * KP; e += nt) {
            wD[e] = 0.0; wRq[e] = 0.0; wVtV[e] = 0.0; hT[e] = 0.0;
            hVc[e] = ((e / KP) == (e % KP)) ? 1.0 : 0.0;
        }
        if (tid == 0) sFlag = 1;
        __syncthreads();
        if (jeff > 0) {
            // ---- Gram blocks: S'Y = A2' A1, Y' diag(alpha) Y = A1' A1 (one warp per entry, lanes over rows) ----
            for (int e = warp; e < 2 * jeff * jeff; e += KG_WARPS) {
                const int which = e / (jeff * jeff), r = e % (jeff * jeff), a = r / jeff, b = r % jeff;
                const int ca = which == 0 ? jeff + a : a;
                double acc = 0.0;
                for (int i = lane; i < n; i += 32) acc = fma(PNL(i, ca), PNL(i, b), acc);
                acc = kg_warp_sum(acc);
                if (lane == 0) (which == 0 ? wStY : wYaY)[a * JM + b] = acc;
            }
            __syncthreads();
            // ---- D (src/inverse_hessian.jl:119-130): nRinv = -triu(S'Y)^-1 by back substitution ----------
            if (tid < jeff) {
                const int b = tid;
                for (int a = jeff - 1; a >= 0; --a) {
                    double rhs = (a == b) ? -1.0 : 0.0;
                    for (int c = a + 1; c < jeff; ++c) rhs -= wStY[a * JM + c] * wNR[c * JM + b];
                    wNR[a * JM + b] = (a <= b) ? rhs / wStY[a * JM + a] : 0.0;
                }
            }
            __syncthreads();
            for (int e = tid; e < jeff * jeff; e += nt) {
                const int a = e / jeff, b = e % jeff, lo = min(a, b), hi = max(a, b);
                wM[a * JM + b] = wYaY[lo * JM + hi] + ((a == b) ? wStY[a * JM + a] : 0.0);
            }
            __syncthreads();
            for (int e = tid; e < jeff * jeff; e += nt) {
                const int a = e / jeff, b = e % jeff;
                double s = 0.0;
                for (int c = 0; c <= b; ++c) s = fma(wM[a * JM + c], wNR[c * JM + b], s);
                wE[a * KP + b] = s;
            }
            __syncthreads();
            for (int e = tid; e < jeff * jeff; e += nt) {
                const int a = e / jeff, b = e % jeff;
                double s = 0.0;
                for (int c = 0; c <= a; ++c) s = fma(wNR[c * JM + a], wE[c * KP + b], s);
                wD[(jeff + a) * KP + jeff + b] = s;
                wD[a * KP + jeff + b] = wNR[a * JM + b];
                wD[(jeff + a) * KP + b] = wNR[b * JM + a];
            }
            __syncthreads();
            // ---- Householder QR in place (dgeqr2 / dlarfg / dlarf) -------------------------------------
            for (int j = 0; j < kq; ++j) {
                // dots of the unscaled column j (rows i > j) with itself and with every other column
                for (int c = warp; c < kc; c += KG_WARPS) {
                    double acc = 0.0;
                    for (int i = j + 1 + lane; i < n; i += 32) acc = fma(PNL(i, j), PNL(i, c), acc);
                    acc = kg_warp_sum(acc);
                    if (lane == 0) wDots[c] = acc;
                }
                __syncthreads();
                const double ss = wDots[j], ajj = PNL(j, j);
                double tau = 0.0, beta = ajj, scale = 0.0;
                if (ss != 0.0) {
                    beta = -copysign(sqrt(fma(ajj, ajj, ss)), ajj);
                    tau = (beta - ajj) / beta;
                    scale = 1.0 / (ajj - beta);
                }
                __syncthreads();  // everybody has read wDots[j], PNL(j, j)
                // v'a_c for c != j (row j contributes a_jc because v_j = 1); 0 when tau == 0
                for (int c = tid; c < kc; c += nt)
                    if (c != j) wDots[c] = (tau != 0.0) ? fma(scale, wDots[c], PNL(j, c)) : 0.0;
                __syncthreads();
                if (tau != 0.0) {
                    for (int i = j + tid; i < n; i += nt) {
                        if (i > j) {
                            const double v = PNL(i, j) * scale;
                            PNL(i, j) = v;
                            for (int c = j + 1; c < kc; ++c) PNL(i, c) = fma(v, -tau * wDots[c], PNL(i, c));
                        } else {
                            for (int c = j + 1; c < kc; ++c) PNL(i, c) = PNL(i, c) - tau * wDots[c];
                            PNL(i, j) = beta;
                        }
                    }
                }
                if (tid == 0) wTau[j] = tau;
                for (int c = tid; c < j; c += nt) wVtV[c * KP + j] = wDots[c];
                __syncthreads();
            }
            // ---- T (dlarft, forward columnwise) --------------------------------------------------------
            if (warp == 0) {
                for (int j = 0; j < kq; ++j) {
                    const double tau = wTau[j];
                    for (int a = lane; a < j; a += 32) {
                        double tv = 0.0;
                        for (int b = a; b < j; ++b) tv = fma(hT[a * KP + b], wVtV[b * KP + j], tv);
                        wW[a] = -tau * tv;
                    }
                    __syncwarp();
                    for (int a = lane; a < j; a += 32) hT[a * KP + j] = wW[a];
                    if (lane == 0) hT[j * KP + j] = tau;
                    __syncwarp();
                }
            }
            __syncthreads();
            // ---- Rq out, Vh fixed up (unit diagonal, zeros above and in unused columns) -----------------
            for (int i = tid; i < n; i += nt) {
                if (i < kq) {
                    for (int c = i; c < kc; ++c) wRq[i * KP + c] = PNL(i, c);
                    PNL(i, i) = 1.0;
                    for (int c = i + 1; c < KP; ++c) PNL(i, c) = 0.0;
                } else {
                    for (int c = kq; c < kc; ++c) PNL(i, c) = 0.0;
                }
            }
            __syncthreads();
            // E = D Rq' (kc x kq), C = I + Rq E (kq x kq)
            for (int e = tid; e < kc * kq; e += nt) {
                const int c = e / kq, b = e % kq;
                double s = 0.0;
                for (int d = b; d < kc; ++d) s = fma(wD[c * KP + d], wRq[b * KP + d], s);
                wE[c * KP + b] = s;
            }
            __syncthreads();
            for (int e = tid; e < kq * kq; e += nt) {
                const int a = e / kq, b = e % kq;
                double s = 0.0;
                for (int c = a; c < kc; ++c) s = fma(wRq[a * KP + c], wE[c * KP + b], s);
                wC[a * KP + b] = s + ((a == b) ? 1.0 : 0.0);
            }
            __syncthreads();
            // Cholesky C = Vc' Vc (upper), dpotrf semantics
            if (warp == 0) {
                for (int j = 0; j < kq; ++j) {
                    double d = wC[j * KP + j];
                    for (int m = 0; m < j; ++m) d = fma(-hVc[m * KP + j], hVc[m * KP + j], d);
                    const bool ok = (d > 0.0);
                    const double vjj = sqrt(d);
                    __syncwarp();
                    if (lane == 0) {
                        hVc[j * KP + j] = ok ? vjj : NAN;
                        if (!ok) sFlag = 0;
                    }
                    for (int b = j + 1 + lane; b < kq; b += 32) {
                        double s = wC[j * KP + b];
                        for (int m = 0; m < j; ++m) s = fma(-hVc[m * KP + j], hVc[m * KP + b], s);
                        hVc[j * KP + b] = ok ? s / vjj : NAN;
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        // ---- logdet = 2 (logdet U + logdet V) ---------------------------------------------------------------
        {
            double acc = 0.0;
            for (int i = tid; i < n; i += nt) acc += log(PNL(i, KP));
            acc = kg_warp_sum(acc);
            __shared__ double sW[KG_WARPS];
            if (lane == 0) sW[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int w = 0; w < KG_WARPS; ++w) s += sW[w];
                for (int j = 0; j < kq; ++j) s += log(hVc[j * KP + j]);
                sB[0] = 2.0 * s;
            }
            __syncthreads();
        }
        // ---- mu = theta + L (R g), t = U g already in the mu slot --------------------------------------------
        if (kq > 0) {
            for (int pass = 0; pass < 2; ++pass) {
                for (int c = warp; c < kq; c += KG_WARPS) {  // w = Vh' t
                    double acc = 0.0;
                    for (int i = lane; i < n; i += 32) acc = fma(PNL(i, c), PNL(i, KP + 1), acc);
                    acc = kg_warp_sum(acc);
                    if (lane == 0) wW2[c] = acc;
                }
                __syncthreads();
                for (int a = tid; a < kq; a += nt) {  // pass 0: T' w;  pass 1: T w
                    double s = 0.0;
                    if (pass == 0) {
                        for (int c = 0; c <= a; ++c) s = fma(hT[c * KP + a], wW2[c], s);
                    } else {
                        for (int c = a; c < kq; ++c) s = fma(hT[a * KP + c], wW2[c], s);
                    }
                    wW[a] = s;
                }
                __syncthreads();
                for (int i = tid; i < n; i += nt) {
                    double t = PNL(i, KP + 1);
                    for (int c = 0; c < kq; ++c) t = fma(-PNL(i, c), wW[c], t);
                    PNL(i, KP + 1) = t;
                }
                __syncthreads();
                if (pass == 0) {  // head <- Vc' (Vc head)
                    for (int a = tid; a < kq; a += nt) {
                        double s = 0.0;
                        for (int c = a; c < kq; ++c) s = fma(hVc[a * KP + c], PNL(c, KP + 1), s);
                        wW2[a] = s;
                    }
                    __syncthreads();
                    for (int a = tid; a < kq; a += nt) {
                        double s = 0.0;
                        for (int c = 0; c <= a; ++c) s = fma(hVc[c * KP + a], wW2[c], s);
                        PNL(a, KP + 1) = s;
                    }
                    __syncthreads();
                }
            }
        }
        for (int i = tid; i < n; i += nt) PNL(i, KP + 1) = fma(PNL(i, KP), PNL(i, KP + 1), theta[i]);
        if (tid == 0) {
            hdr[PFB_HDR_LOGDET(KP)] = sB[0];
            hdr[PFB_HDR_FLAG(KP)] = (double)sFlag;
            hdr[PFB_HDR_KEFF(KP)] = (double)kq;
        }
        __syncthreads();
#undef PNL
    }
}

// ---- K3g ------------------------------------------------------------------------------------------------
struct kg_sel {
    const int32_t* cnt;
    const int2* list;
    int cap;
};

// One CTA per slot, one warp per draw: u (contract normals, or host normals), u~ = diag(Vc', I) u,
// z = u~ - Vh (T (Vh' u~)), x = mu + sqrt(alpha) .* z; log q; log p of the diagonal-quadratic families.
// The draw lives in `xbuf` (the draws output, or a per-warp scratch row when no draws are wanted).
__global__ void __launch_bounds__(KG_THREADS)
pfb_k3g_elbo_sample(int KP, int model, int n, int K, const int32_t* __restrict__ unit_list,
                    const double* __restrict__ FRg, const double* __restrict__ HDR, const uint64_t* __restrict__ seeds,
                    const double* __restrict__ u_host, const double* __restrict__ mp0, const double* __restrict__ mp1,
                    double mc0, double* __restrict__ logp_out, double* __restrict__ logq_out,
                    double* __restrict__ draws_out, kg_sel sel, const double* __restrict__ fbX,
                    const double* __restrict__ fbG, const int64_t* __restrict__ fb_off,
                    const uint64_t* __restrict__ fb_seeds, const int32_t* __restrict__ fb_path,
                    double* __restrict__ scratch /* [grid][KG_WARPS][n + 2 KP] */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = blockIdx.x;
    const int unit = unit_list ? unit_list[slot] : slot;
    const bool SEL = sel.cnt != nullptr;
    const int2* sel_list = SEL ? sel.list + (int64_t)slot * sel.cap : nullptr;
    const int Kslot = SEL ? sel.cnt[slot] : K;
    const int RS = KP + 2;
    double* wrow = scratch + ((size_t)blockIdx.x * KG_WARPS + warp) * (size_t)(n + 2 * KP);
    double* wvec = wrow + n;       // w = Vh' u~, then c = T w   (KP)
    double* cvec = wvec + KP;      // (KP)
    const bool fallback = unit < 0 && fbX != nullptr;
    if (unit < 0 && !fallback) {
        for (int j = warp; j < Kslot; j += KG_WARPS) {
            const int64_t oc = SEL ? (int64_t)sel_list[j].y : (int64_t)slot * K + j;
            if (!SEL && lane == 0) {
                if (logp_out) logp_out[oc] = NAN;
                if (logq_out) logq_out[oc] = NAN;
            }
            if (draws_out)
                for (int i = lane; i < n; i += 32) draws_out[oc * n + i] = NAN;
        }
        return;
    }
    const double* fr = fallback ? nullptr : FRg + (int64_t)unit * n * RS;
    const double* hdr = fallback ? nullptr : HDR + (int64_t)unit * pfb_hs_of(KP);
    const int path = fallback ? (fb_path ? fb_path[slot] : slot) : 0;
    const double* x0 = fallback ? fbX + fb_off[path] * (int64_t)n : nullptr;
    const double* g0 = fallback ? fbG + fb_off[path] * (int64_t)n : nullptr;
    const uint64_t seed = fallback ? fb_seeds[path] : seeds[unit];
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int kq = fallback ? 0 : (int)hdr[PFB_HDR_KEFF(KP)];
    const double logdet = fallback ? 0.0 : hdr[PFB_HDR_LOGDET(KP)];
    const bool pd_ok = fallback ? true : (hdr[PFB_HDR_FLAG(KP)] != 0.0);
    const double* hT = hdr;
    const double* hVc = fallback ? nullptr : hdr + KP * KP;
    for (int j = warp; j < Kslot; j += KG_WARPS) {
        const uint32_t kdraw = SEL ? (uint32_t)sel_list[j].x : (uint32_t)j;
        const int64_t oc = SEL ? (int64_t)sel_list[j].y : (int64_t)slot * K + j;
        double* x = draws_out ? draws_out + oc * n : wrow;
        // normals (row pairs over the lanes) and |u|^2
        double usq = 0.0;
        for (int rp = lane; 2 * rp < n; rp += 32) {
            double z0, z1 = 0.0;
            const int i = 2 * rp;
            if (u_host != nullptr && !fallback) {
                const double* uh = u_host + ((int64_t)unit * K + kdraw) * n;
                z0 = uh[i];
                if (i + 1 < n) z1 = uh[i + 1];
            } else {
                pf_normal_pair((uint32_t)rp, kdraw, k0, k1, PF_ZIG_XK_DEV, PF_ZIG_F_DEV, &z0, &z1);
            }
            x[i] = z0;
            usq = fma(z0, z0, usq);
            if (i + 1 < n) {
                x[i + 1] = z1;
                usq = fma(z1, z1, usq);
            }
        }
        usq = kg_warp_sum(usq);
        __syncwarp();
        if (kq > 0) {
            // head: u~[0..kq) = Vc' u[0..kq)   (src/woodbury.jl:139)
            for (int a = lane; a < kq; a += 32) {
                double s = 0.0;
                for (int m = 0; m <= a; ++m) s = fma(hVc[m * KP + a], x[m], s);
                cvec[a] = s;
            }
            __syncwarp();
            for (int a = lane; a < kq; a += 32) x[a] = cvec[a];
            __syncwarp();
            // w = Vh' u~
            for (int c = 0; c < kq; ++c) {
                double acc = 0.0;
                for (int i = lane; i < n; i += 32) acc = fma(fr[(int64_t)i * RS + c], x[i], acc);
                acc = kg_warp_sum(acc);
                if (lane == 0) wvec[c] = acc;
            }
            __syncwarp();
            for (int a = lane; a < kq; a += 32) {  // c = T w
                double s = 0.0;
                for (int b = a; b < kq; ++b) s = fma(hT[a * KP + b], wvec[b], s);
                cvec[a] = s;
            }
            __syncwarp();
        }
        // x = mu + sqrt(alpha) (u~ - Vh c), model sums
        double ma = 0.0, mb = 0.0;
        for (int i = lane; i < n; i += 32) {
            double xi;
            if (fallback) {
                xi = (x0[i] + g0[i]) + x[i];
            } else {
                double z = x[i];
                for (int c = 0; c < kq; ++c) z = fma(-fr[(int64_t)i * RS + c], cvec[c], z);
                xi = fma(fr[(int64_t)i * RS + KP], z, fr[(int64_t)i * RS + KP + 1]);
            }
            if (draws_out) x[i] = xi;
            if (model == PFB_MODEL_ISONORMAL) {
                ma = fma(xi, xi, ma);
            } else if (model == PFB_MODEL_FUNNEL) {
                if (i == 0) mb = xi; else ma = fma(xi, xi, ma);
            } else if (model == PFB_MODEL_DIAGNORMAL) {
                const double zz = (xi - mp0[i]) * mp1[i];
                ma = fma(zz, zz, ma);
            }
        }
        ma = kg_warp_sum(ma);
        mb = kg_warp_sum(mb);
        if (lane == 0 && !SEL) {
            double lp;
            if (model == PFB_MODEL_ISONORMAL) lp = ma / -2.0;
            else if (model == PFB_MODEL_FUNNEL) {
                const double t3 = mb / 3.0;
                lp = (fma(t3, t3, (double)(n - 1) * mb) + exp(-mb) * ma) / -2.0;
            } else if (model == PFB_MODEL_DIAGNORMAL) lp = fma(ma, -0.5, mc0);
            else lp = 0.0;  // GEMM-shaped / host targets: filled in by K8g / the callback
            double lq = (fma((double)n, PFB_LOG2PI, logdet) + usq) / -2.0;
            if (!pd_ok) lq = NAN;
            if (logp_out) logp_out[oc] = lp;
            if (logq_out) logq_out[oc] = lq;
        }
        __syncwarp();
    }
}

// gather of fitted normals (units list; < 0: NaN) from the generic layout
__global__ void pfb_kg_gather_fit(int n, int KP, const int32_t* __restrict__ units, const double* __restrict__ FRg,
                                  const double* __restrict__ HDR, const double* __restrict__ alpha,
                                  const int32_t* __restrict__ hist_cnt, double* mu, double* al, double* vh, double* Tm,
                                  double* Vc, double* logdet, int32_t* jeff) {
    const int p = blockIdx.x;
    const int u = units[p];
    const int RS = KP + 2, HS = pfb_hs_of(KP);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double* row = u >= 0 ? FRg + ((int64_t)u * n + i) * RS : nullptr;
        mu[(int64_t)p * n + i] = u >= 0 ? row[KP + 1] : NAN;
        al[(int64_t)p * n + i] = u >= 0 ? alpha[(int64_t)u * n + i] : NAN;
        for (int j = 0; j < KP; ++j) vh[((int64_t)p * KP + j) * n + i] = u >= 0 ? row[j] : NAN;
    }
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
        Tm[(int64_t)p * KP * KP + e] = u >= 0 ? HDR[(int64_t)u * HS + e] : NAN;
        Vc[(int64_t)p * KP * KP + e] = u >= 0 ? HDR[(int64_t)u * HS + KP * KP + e] : NAN;
    }
    if (threadIdx.x == 0) {
        logdet[p] = u >= 0 ? HDR[(int64_t)u * HS + 2 * KP * KP] : NAN;
        jeff[p] = u >= 0 ? hist_cnt[u] : 0;
    }
}

extern "C" size_t pfb_kg_k2_workspace_doubles(int KP, int grid) { return kg_ws_doubles(KP) * (size_t)grid; }
extern "C" int pfb_kg_k2_grid(int U) { return U < 296 ? U : 296; }
extern "C" cudaError_t pfb_launch_k2g(cudaStream_t st, int KP, int n, int u_base, int U, int J, const double* X,
                                      const double* G, const int32_t* unit_col, const double* alpha,
                                      const int32_t* hist, const int32_t* hist_cnt, double* FRg, double* HDR,
                                      double* WS) {
    if (U <= 0) return cudaSuccess;
    pfb_k2g_woodbury_build<<<pfb_kg_k2_grid(U), KG_THREADS, 0, st>>>(KP, n, U, u_base, J, X, G, unit_col, alpha, hist,
                                                                     hist_cnt, FRg, HDR, WS);
    return cudaGetLastError();
}
extern "C" size_t pfb_kg_k3_scratch_doubles(int KP, int n, int nslots) {
    return (size_t)nslots * KG_WARPS * (size_t)(n + 2 * KP);
}
extern "C" cudaError_t pfb_launch_k3g(cudaStream_t st, int KP, int model, int n, int K, int nslots,
                                      const int32_t* unit_list, const double* FRg, const double* HDR,
                                      const uint64_t* seeds, const double* u_host, const double* mp0, const double* mp1,
                                      double mc0, double* logp, double* logq, double* draws, const int32_t* sel_cnt,
                                      const void* sel_list, int sel_cap, const double* fbX, const double* fbG,
                                      const int64_t* fb_off, const uint64_t* fb_seeds, const int32_t* fb_path,
                                      double* scratch) {
    if (nslots <= 0) return cudaSuccess;
    if (sel_cnt != nullptr && draws == nullptr) return cudaErrorInvalidValue;
    kg_sel sel{sel_cnt, reinterpret_cast<const int2*>(sel_list), sel_cap};
    pfb_k3g_elbo_sample<<<nslots, KG_THREADS, 0, st>>>(KP, model, n, K, unit_list, FRg, HDR, seeds, u_host, mp0, mp1, mc0,
                                                       logp, logq, draws, sel, fbX, fbG, fb_off, fb_seeds, fb_path,
                                                       scratch);
    return cudaGetLastError();
}
extern "C" cudaError_t pfb_launch_kg_gather_fit(cudaStream_t st, int P, int n, int KP, const int32_t* units,
                                                const double* FRg, const double* HDR, const double* alpha,
                                                const int32_t* hist_cnt, double* mu, double* al, double* vh, double* Tm,
                                                double* Vc, double* logdet, int32_t* jeff) {
    if (P <= 0) return cudaSuccess;
    pfb_kg_gather_fit<<<P, 256, 0, st>>>(n, KP, units, FRg, HDR, alpha, hist_cnt, mu, al, vh, Tm, Vc, logdet, jeff);
    return cudaGetLastError();
}
