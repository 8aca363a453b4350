// K2 woodbury_build — one CTA per unit (path p, iteration l >= 1).
//
// Replaces, for H0 = Diagonal(alpha):
//   lbfgs_inverse_hessian   (reference: src/inverse_hessian.jl:98-133)   B = [alpha.*Y | S], D
//   pdfactorize             (reference: src/woodbury.jl:201-207)          U = sqrt(alpha),
//                           Q,R = qr(U' \ B) (Householder, LAPACK dlarfg/dlarft conventions so
//                           that the full n x n Q equals the reference's), V = chol(I + R D R')
//   logabsdet               (reference: src/woodbury.jl:77-80)
//   mu = Sigma * g + theta  (reference: src/mvnormal.jl:17 -> src/woodbury.jl:346-349, :64-68)
//
// Output: the unit's factor record rows { Vh[i][0..KP), sqrt(alpha_i), mu_i } and header
// { T, Vc, logdet, pd flag, k_eff } (layout: pfb_common.cuh).  The QR runs in place in the
// record rows (global memory, L1/L2 resident: n * (KP+2) * 8 bytes per unit); every thread owns
// whole rows, so the only cross-thread traffic is the block reductions.
#include "pfb_common.cuh"

#define PFB_K2_THREADS 256
#define PFB_K2_THREADS_SMEM 512
#define PFB_K2_THREADS_REG 256  // REGCOLS variant: two CTAs per SM
#define PFB_K2_RPT 4            // rows per thread whose sqrt(alpha) / mu entries live in registers (n <= 1024)
#define PFB_K2_SCRATCH(KP) ((KP) * 16 * 8 + (KP) + 1)  // doubles; keeps the panel 16-byte aligned for even KP
#define PFB_K2_SCRATCH_REG(KP) ((KP) * 8 * 8 + (KP) + 1)  // the same for the 8 warps of the REGCOLS variant

// The unit's n x (KP+2) panel { A~ -> Vh, sqrt(alpha), t -> mu } lives in shared memory, column-major
// (conflict-free for thread-per-row access), when it fits (SMEM_PANEL: n (KP+2) 8 bytes <= ~200 KB,
// i.e. n <= 1800 at history 6); otherwise in the FR rows in global memory (L2 resident).
// REGCOLS (n <= 1024, 256 threads): the two extra columns { sqrt(alpha), t -> mu } stay in the registers
// of the thread that owns the row (4 rows per thread), which shrinks the panel to n x KP — 98 KB at
// n = 1024, KP = 12 — so that TWO CTAs share an SM and one unit's block-wide reductions (a dozen
// Householder steps, each a barrier) overlap the other's arithmetic.
template <int KP, bool SMEM_PANEL, bool REGCOLS = false>
__global__ void __launch_bounds__(REGCOLS ? PFB_K2_THREADS_REG : (SMEM_PANEL ? PFB_K2_THREADS_SMEM : PFB_K2_THREADS),
                                  REGCOLS ? 2 : 1)
pfb_k2_woodbury_build(int n, int J, const double* __restrict__ X, const double* __restrict__ G,
                      const int32_t* __restrict__ unit_col, const double* __restrict__ alpha_all,
                      const int32_t* __restrict__ hist, const int32_t* __restrict__ hist_cnt,
                      double* __restrict__ FR, double* __restrict__ HDR, double* __restrict__ FR2,
                      int model, const double* __restrict__ mp0, const double* __restrict__ mp1, int u_base) {
    constexpr int RS = KP + 2;
    constexpr int JM = KP / 2;
    __shared__ double sStY[JM][JM], sYaY[JM][JM], sNRinv[JM][JM], sM[JM][JM];
    __shared__ double sD[KP][KP], sRq[KP][KP], sE[KP][KP], sC[KP][KP], sVc[KP][KP], sT[KP][KP];
    __shared__ double sVtV[KP][KP];
    __shared__ double sTau[KP], sHead[KP], sW[KP], sW2[KP];
    __shared__ double sBcast[2];
    __shared__ int sFlag;

    const int u = blockIdx.x + u_base;  // u_base: first unit of this launch (pipelined uploads)
    const int tid = threadIdx.x, nt = blockDim.x;
    const int jeff = hist_cnt[u];
    const int kc = 2 * jeff;
    const int kq = min(n, kc);
    extern __shared__ __align__(16) double s_dyn[];  // scratch of pfb_block_sum_fast, then the panel
    double* scratch = s_dyn;                          // KP * 16 * 8 + KP (<= 16 warps)
    double* s_panel = s_dyn + (REGCOLS ? PFB_K2_SCRATCH_REG(KP) : PFB_K2_SCRATCH(KP));  // SMEM_PANEL: [KP+2][ldp] ([KP][ldp] under REGCOLS)
    const int ldp = (n + 1) | 1;  // odd leading dimension
    double* fr = SMEM_PANEL ? nullptr : FR + (int64_t)u * n * RS;
    // element (row i, column c) of the panel
#define PNL(i, c) (*(SMEM_PANEL ? (s_panel + (c) * ldp + (i)) : (fr + (int64_t)(i) * RS + (c))))
    double* hdr = HDR + (int64_t)u * pfb_hs_of(KP);
    // the two extra columns: registers (REGCOLS) or panel columns KP, KP + 1
    double rsa[REGCOLS ? PFB_K2_RPT : 1], rtm[REGCOLS ? PFB_K2_RPT : 1];
#define XSA(r, i) (*(REGCOLS ? &rsa[r] : &PNL(i, KP)))
#define XTM(r, i) (*(REGCOLS ? &rtm[r] : &PNL(i, KP + 1)))
    // f(r, i) for every row i this thread owns (r = its register slot under REGCOLS)
    auto each_row = [&](auto&& f) {
        if constexpr (REGCOLS) {
#pragma unroll
            for (int r = 0; r < PFB_K2_RPT; ++r) {
                const int i = tid + r * PFB_K2_THREADS_REG;
                if (i < n) f(r, i);
            }
        } else {
            for (int i = tid; i < n; i += nt) f(0, i);
        }
    };

    // Weighted Gram matrix of the panel's first KP columns on the FP64 tensor cores:
    //   Gm[a][b] = sum_i wgt(i) P(i, a) P(i, b),  a, b < KP  (row-major, KP x KP, shared memory).
    // Each warp takes a slice of rows (4 per DMMA k-step); A[m = col][k = row] and B[k = row][n = col]
    // are the same panel element, B scaled by the weight.  Warp partials are summed in warp order
    // through `scratch` (one 8 x 8 tile at a time) => deterministic.
    constexpr int GT = (KP + 7) / 8;
    auto gram = [&](auto wgt, double* Gm) {
        const int lane = tid & 31, warp = tid >> 5, nwarp = (nt + 31) >> 5;
        const int g = lane >> 2, t = lane & 3;
        double acc[GT][GT][2];
#pragma unroll
        for (int ta = 0; ta < GT; ++ta)
#pragma unroll
            for (int tb = 0; tb < GT; ++tb) acc[ta][tb][0] = acc[ta][tb][1] = 0.0;
        const int ksteps = (n + 3) >> 2;
        for (int ks = warp; ks < ksteps; ks += nwarp) {
            const int row = 4 * ks + t;
            const bool rok = row < n;
            const double w = rok ? wgt(row) : 0.0;
            double fa[GT], fb[GT];
#pragma unroll
            for (int ta = 0; ta < GT; ++ta) {
                const int col = 8 * ta + g;
                fa[ta] = (rok && col < KP) ? PNL(row, col) : 0.0;
                fb[ta] = fa[ta] * w;
            }
#pragma unroll
            for (int ta = 0; ta < GT; ++ta)
#pragma unroll
                for (int tb = 0; tb < GT; ++tb)
                    if (tb >= ta)  // upper block triangle; mirrored below
                        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                            : "+d"(acc[ta][tb][0]), "+d"(acc[ta][tb][1])
                            : "d"(fa[ta]), "d"(fb[tb]));
        }
#pragma unroll
        for (int ta = 0; ta < GT; ++ta)
#pragma unroll
            for (int tb = 0; tb < GT; ++tb) {
                if (tb < ta) continue;
                __syncthreads();  // scratch free
                // D fragment: rows 8 ta + g, columns 8 tb + 2 t + {0, 1}
                scratch[(warp * 8 + g) * 8 + 2 * t] = acc[ta][tb][0];
                scratch[(warp * 8 + g) * 8 + 2 * t + 1] = acc[ta][tb][1];
                __syncthreads();
                if (tid < 64) {
                    double sum = 0.0;
                    for (int wq = 0; wq < nwarp; ++wq) sum += scratch[wq * 64 + tid];
                    const int a = 8 * ta + (tid >> 3), b = 8 * tb + (tid & 7);
                    // diagonal tiles: (a, b) and (b, a) are both computed but round differently when the
                    // weight is not 1 (a (b w) vs b (a w)) — only the upper triangle writes, so the result
                    // is symmetric and deterministic
                    if (a < KP && b < KP && (ta != tb || a <= b)) {
                        Gm[a * KP + b] = sum;
                        Gm[b * KP + a] = sum;
                    }
                }
            }
        __syncthreads();
    };
    const double* alpha = alpha_all + (int64_t)u * n;
    const int64_t col = unit_col[u];
    const double* theta = X + col * n;
    const double* g = G + col * n;
    const int32_t* hu = hist + (int64_t)u * J;

    // ---- Phase A: A~ = U' \ B = [ (alpha .* Y) ./ sqrt(alpha) | S ./ sqrt(alpha) ] -------------
    each_row([&](int r, int i) {
        double a = alpha[i];
        double sa = sqrt(a);
        for (int j = 0; j < jeff; ++j) {
            int64_t c = hu[j];
            double s = X[(c + 1) * n + i] - X[c * n + i];
            double y = G[c * n + i] - G[(c + 1) * n + i];
            PNL(i, j) = (a * y) / sa;
            PNL(i, jeff + j) = s / sa;
        }
        for (int j = kc; j < KP; ++j) PNL(i, j) = 0.0;
        XSA(r, i) = sa;
        XTM(r, i) = sa * g[i];  // scratch for phase H: t = U g
    });
    if (tid == 0) sFlag = 1;
    // zero-init small matrices
    for (int e = tid; e < KP * KP; e += nt) {
        (&sD[0][0])[e] = 0.0;
        (&sT[0][0])[e] = 0.0;
        (&sRq[0][0])[e] = 0.0;
        (&sVtV[0][0])[e] = 0.0;
        (&sVc[0][0])[e] = ((e / KP) == (e % KP)) ? 1.0 : 0.0;
    }
    __syncthreads();

    if (jeff > 0) {
        // ---- Gram blocks of A~ (DMMA): S'Y = A2' A1, Y' diag(alpha) Y = A1' A1 ---------------------
        gram([](int) { return 1.0; }, &sE[0][0]);
        for (int e = tid; e < jeff * jeff; e += nt) {
            const int a = e / jeff, b = e % jeff;
            sStY[a][b] = sE[jeff + a][b];
            sYaY[a][b] = sE[a][b];
        }
        __syncthreads();
        // ---- D  (src/inverse_hessian.jl:119-130) ----------------------------------------------
        // R = triu(S'Y); nRinv = -R^-1 (back substitution, one column per thread)
        if (tid < jeff) {
            const int b = tid;
            for (int a = jeff - 1; a >= 0; --a) {
                double rhs = (a == b) ? -1.0 : 0.0;
                for (int c = a + 1; c < jeff; ++c) rhs -= sStY[a][c] * sNRinv[c][b];
                sNRinv[a][b] = (a <= b) ? rhs / sStY[a][a] : 0.0;
            }
        }
        __syncthreads();
        // M = Diagonal(R) + Y' H0 Y, symmetrised from its upper triangle (copytri! 'U')
        for (int e = tid; e < jeff * jeff; e += nt) {
            int a = e / jeff, b = e % jeff;
            int lo = min(a, b), hi = max(a, b);
            sM[a][b] = sYaY[lo][hi] + ((a == b) ? sStY[a][a] : 0.0);
        }
        __syncthreads();
        // tmp = M * nRinv (rmul!), D22 = nRinv' * tmp (lmul!)
        for (int e = tid; e < jeff * jeff; e += nt) {
            int a = e / jeff, b = e % jeff;
            double s = 0.0;
            for (int c = 0; c <= b; ++c) s = fma(sM[a][c], sNRinv[c][b], s);
            sE[a][b] = s;
        }
        __syncthreads();
        for (int e = tid; e < jeff * jeff; e += nt) {
            int a = e / jeff, b = e % jeff;
            double s = 0.0;
            for (int c = 0; c <= a; ++c) s = fma(sNRinv[c][a], sE[c][b], s);
            sD[jeff + a][jeff + b] = s;
            sD[a][jeff + b] = sNRinv[a][b];
            sD[jeff + a][b] = sNRinv[b][a];
        }
        __syncthreads();

        // ---- Phase C: Householder QR in place (dgeqr2 / dlarfg / dlarf conventions) ------------
        for (int j = 0; j < kq; ++j) {
            // one reduction per reflector: acc[j] = sum_{i>j} a_ij^2, acc[c] = sum_{i>j} a_ij a_ic on
            // the unscaled column; v = scale * a_j (v_j = 1) then gives v'a_c = scale acc[c] + a_jc
            double acc[KP];
#pragma unroll
            for (int c = 0; c < KP; ++c) acc[c] = 0.0;
            for (int i = tid; i < n; i += nt) {
                if (i > j) {
                    const double x = PNL(i, j);
#pragma unroll
                    for (int c = 0; c < KP; ++c)
                        if (c < kc) acc[c] = fma(x, (c == j) ? x : PNL(i, c), acc[c]);
                }
            }
            pfb_block_sum_fast<KP>(acc, scratch, (kc >= 32) ? 0xffffffffu : ((1u << kc) - 1u));
            double ss = 0.0;
#pragma unroll
            for (int c = 0; c < KP; ++c)
                if (c == j) ss = acc[c];
            const double ajj = PNL(j, j);
            double tau = 0.0, beta = ajj, scale = 0.0;
            if (ss != 0.0) {
                beta = -copysign(sqrt(fma(ajj, ajj, ss)), ajj);
                tau = (beta - ajj) / beta;
                scale = 1.0 / (ajj - beta);
            }
            // v'a_c for c != j (row j contributes a_jc because v_j = 1); 0 when tau == 0
#pragma unroll
            for (int c = 0; c < KP; ++c)
                acc[c] = (tau != 0.0 && c < kc && c != j) ? fma(scale, acc[c], PNL(j, c)) : 0.0;
            __syncthreads();  // every thread has read row j before it is updated
            if (tau != 0.0) {
                for (int i = tid; i < n; i += nt) {
                    if (i > j) {
                        const double v = PNL(i, j) * scale;
                        PNL(i, j) = v;
#pragma unroll
                        for (int c = 0; c < KP; ++c)
                            if (c > j && c < kc) PNL(i, c) = fma(v, -tau * acc[c], PNL(i, c));
                    } else if (i == j) {
#pragma unroll
                        for (int c = 0; c < KP; ++c)
                            if (c > j && c < kc) PNL(i, c) = PNL(i, c) - tau * acc[c];
                        PNL(i, j) = beta;
                    }
                }
            }
            if (tid == 0) {
                sTau[j] = tau;
#pragma unroll
                for (int c = 0; c < KP; ++c)
                    if (c < j) sVtV[c][j] = acc[c];  // V(:,c)' v_j for c < j  (0 when tau == 0)
            }
            __syncthreads();
        }

        // ---- Phase D: T (dlarft, forward columnwise), lane a owns row a -------------------------
        if (tid < 32) {
            for (int j = 0; j < kq; ++j) {
                const double tau = sTau[j];
                double tv = 0.0;
                if (tid < j) {
                    for (int b = tid; b < j; ++b) tv = fma(sT[tid][b], sVtV[b][j], tv);
                    tv = -tau * tv;
                }
                __syncwarp();
                if (tid < j) sT[tid][j] = tv;
                if (tid == j) sT[j][j] = tau;
                __syncwarp();
            }
        }
        // ---- Phase E: Rq to smem, fix up Vh (unit diagonal, zeros above / in padded columns) ----
        for (int i = tid; i < n; i += nt) {
            if (i < kq) {
                for (int c = i; c < kc; ++c) sRq[i][c] = PNL(i, c);
                PNL(i, i) = 1.0;
                for (int c = i + 1; c < KP; ++c) PNL(i, c) = 0.0;
            } else {
                for (int c = kq; c < kc; ++c) PNL(i, c) = 0.0;
            }
        }
        __syncthreads();
        // E = D * Rq'  (kc x kq),  C = I + Rq * E  (kq x kq, upper triangle used)
        for (int e = tid; e < kc * kq; e += nt) {
            int c = e / kq, b = e % kq;
            double s = 0.0;
            for (int d = b; d < kc; ++d) s = fma(sD[c][d], sRq[b][d], s);
            sE[c][b] = s;
        }
        __syncthreads();
        for (int e = tid; e < kq * kq; e += nt) {
            int a = e / kq, b = e % kq;
            double s = 0.0;
            for (int c = a; c < kc; ++c) s = fma(sRq[a][c], sE[c][b], s);
            sC[a][b] = s + ((a == b) ? 1.0 : 0.0);
        }
        __syncthreads();
        // Cholesky C = Vc' Vc (upper), dpotrf semantics: fail on pivot <= 0 or NaN
        if (tid < 32) {
            for (int j = 0; j < kq; ++j) {
                double d = sC[j][j];
                for (int m = 0; m < j; ++m) d = fma(-sVc[m][j], sVc[m][j], d);
                const bool ok = (d > 0.0);
                const double vjj = sqrt(d);
                if (tid == 0) {
                    sVc[j][j] = ok ? vjj : NAN;
                    if (!ok) sFlag = 0;
                }
                for (int b = j + 1 + tid; b < kq; b += 32) {
                    double s = sC[j][b];
                    for (int m = 0; m < j; ++m) s = fma(-sVc[m][j], sVc[m][b], s);
                    sVc[j][b] = ok ? s / vjj : NAN;
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }

    // ---- Phase G: logdet = 2 (logdet U + logdet V) ---------------------------------------------
    double ld[1] = {0.0};
    each_row([&](int r, int i) { ld[0] += log(XSA(r, i)); });
    pfb_block_sum_fast<1>(ld, scratch, 1u);
    double ldv = 0.0;
    for (int j = 0; j < kq; ++j) ldv += log(sVc[j][j]);
    const double logdet = 2.0 * (ld[0] + ldv);

    // ---- Phase H: mu = theta + L (R g),  t = U g already in the mu slot ------------------------
    if (kq > 0) {
        for (int pass = 0; pass < 2; ++pass) {
            // w = Vh' t
            double acc[KP];
#pragma unroll
            for (int c = 0; c < KP; ++c) acc[c] = 0.0;
            each_row([&](int r, int i) {
                double t = XTM(r, i);
#pragma unroll
                for (int c = 0; c < KP; ++c)
                    if (c < kq) acc[c] = fma(PNL(i, c), t, acc[c]);
            });
            pfb_block_sum_fast<KP>(acc, scratch, (1u << kq) - 1u);
            if (tid == 0) {
#pragma unroll
                for (int c = 0; c < KP; ++c) sW2[c] = acc[c];
            }
            __syncthreads();
            // pass 0: Q' t = t - Vh (T' w);   pass 1: Q t = t - Vh (T w)
            if (tid < kq) {
                double s = 0.0;
                if (pass == 0) {
                    for (int c = 0; c <= tid; ++c) s = fma(sT[c][tid], sW2[c], s);
                } else {
                    for (int c = tid; c < kq; ++c) s = fma(sT[tid][c], sW2[c], s);
                }
                sW[tid] = s;
            }
            __syncthreads();
            each_row([&](int r, int i) {
                double t = XTM(r, i);
#pragma unroll
                for (int c = 0; c < KP; ++c)
                    if (c < kq) t = fma(-PNL(i, c), sW[c], t);
                XTM(r, i) = t;
                if (pass == 0 && i < kq) sHead[i] = t;
            });
            __syncthreads();
            if (pass == 0) {
                // head <- Vc' (Vc head)      (lmul!(R.V, .) then lmul!(R.V', .))
                if (tid < kq) {
                    double s = 0.0;
                    for (int c = tid; c < kq; ++c) s = fma(sVc[tid][c], sHead[c], s);
                    sW2[tid] = s;
                }
                __syncthreads();
                if (tid < kq) {
                    double s = 0.0;
                    for (int c = 0; c <= tid; ++c) s = fma(sVc[c][tid], sW2[c], s);
                    XTM(0, tid) = s;  // row tid < kq <= KP belongs to thread tid, register slot 0
                }
                __syncthreads();
            }
        }
    }
    // ---- mu, and the diagonal-quadratic statistics of K3's single-pass mode (pfb_common.cuh) --------
    // weights d_i and centres m_i of the registered model family
    auto model_d = [&](int i) -> double {
        if (model == PFB_MODEL_FUNNEL) return i == 0 ? 0.0 : 1.0;
        if (model == PFB_MODEL_DIAGNORMAL) { double is = mp1[i]; return is * is; }  // mp1 = 1 / sd
        return 1.0;
    };
    auto model_m = [&](int i) -> double { return model == PFB_MODEL_DIAGNORMAL ? mp0[i] : 0.0; };
    double e0acc[1] = {0.0};
    each_row([&](int r, int i) {
        // kq == 0: Sigma = diag(alpha): t = sqrt(alpha) g, mu = theta + sqrt(alpha) t
        const double sa = XSA(r, i);
        const double mu = fma(sa, XTM(r, i), theta[i]);
        XTM(r, i) = mu;
        const double d = model_d(i), e = mu - model_m(i);
        e0acc[0] = fma(d * e, e, e0acc[0]);
        if (FR2 != nullptr) {
            // second copy in the tensor-core layout of K3 (pfb_common.cuh): RS2 doubles per row,
            // 4-double groups XOR-swizzled by pfb_swz(row)
            constexpr int RS2 = (KP == 12) ? 16 : 32;
            double* r2 = FR2 + ((int64_t)u * pfb_npad8(n) + i) * RS2;
            const int sw = pfb_swz(i);
            const double pi_ = d * alpha[i], ri_ = 2.0 * (d * sa * e);  // slot KP + 3 holds 2 r (K3 single pass)
#pragma unroll
            for (int c = 0; c < RS2; ++c) {
                double v = (c < KP) ? PNL(i, c)
                                    : (c == KP ? sa : (c == KP + 1 ? mu : (c == KP + 2 ? pi_ : (c == KP + 3 ? ri_ : 0.0))));
                r2[c ^ sw] = v;
            }
        }
    });
    if (FR2 != nullptr) {
        constexpr int RS2 = (KP == 12) ? 16 : 32;
        const int npad = pfb_npad8(n);
        for (int e = n * RS2 + tid; e < npad * RS2; e += nt) FR2[(int64_t)u * npad * RS2 + e] = 0.0;
    }
    pfb_block_sum_fast<1>(e0acc, scratch, 1u);
    // M = Vh' diag(p) Vh on the tensor cores (sE is free by now), rv = Vh' r by one reduction round
    gram([&](int i) { return model_d(i) * alpha[i]; }, &sE[0][0]);
    for (int e = tid; e < KP * KP; e += nt) hdr[PFB_HDR_M(KP) + e] = (&sE[0][0])[e];
    {
        double acc[KP];
#pragma unroll
        for (int b = 0; b < KP; ++b) acc[b] = 0.0;
        each_row([&](int r, int i) {
            const double wgt = model_d(i) * XSA(r, i) * (XTM(r, i) - model_m(i));
#pragma unroll
            for (int b = 0; b < KP; ++b)
                if (b < kq) acc[b] = fma(wgt, PNL(i, b), acc[b]);
        });
        if (kq > 0) pfb_block_sum_fast<KP>(acc, scratch, (1u << kq) - 1u);
        if (tid == 0) {
#pragma unroll
            for (int b = 0; b < KP; ++b) hdr[PFB_HDR_RV(KP) + b] = acc[b];
        }
    }
    // ---- header ---------------------------------------------------------------------------------
    for (int e = tid; e < KP * KP; e += nt) {
        hdr[e] = (&sT[0][0])[e];
        hdr[KP * KP + e] = (&sVc[0][0])[e];
    }
    if (tid == 0) {
        hdr[PFB_HDR_LOGDET(KP)] = logdet;
        hdr[PFB_HDR_FLAG(KP)] = (double)sFlag;
        hdr[PFB_HDR_KEFF(KP)] = (double)kq;
        hdr[PFB_HDR_E0(KP)] = e0acc[0];
    }
}

static size_t k2_panel_bytes(int KP, int n) { return (size_t)(KP + 2) * (size_t)((n + 1) | 1) * 8; }
static size_t k2_scratch_bytes(int KP) { return (size_t)PFB_K2_SCRATCH(KP) * 8; }

// The panel goes to shared memory when it fits next to the kernel's static arrays.
extern "C" int pfb_k2_uses_smem_panel(int KP, int n) {
    int dev = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t stat = (size_t)(9 * KP * KP + 4 * KP + 8) * 8 + 1024;  // static arrays (upper bound)
    return k2_panel_bytes(KP, n) + k2_scratch_bytes(KP) + stat <= (size_t)smem_max;
}

// REGCOLS: two CTAs (dynamic + static shared memory + 1 KB reserved each) must fit one SM
static bool pfb_k2_regcols_fits(int KP, int n) {
    int dev = 0, smem_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    const size_t stat = (size_t)(7 * KP * KP + KP * KP + 4 * KP + 8) * 8 + 256;
    const size_t per_cta = (size_t)KP * (size_t)((n + 1) | 1) * 8 + (size_t)PFB_K2_SCRATCH_REG(KP) * 8 + stat + 1024;
    return 2 * per_cta <= (size_t)smem_sm;
}

template <int KP>
static cudaError_t launch_k2(cudaStream_t st, int n, int u_base, int U, int J, const double* X, const double* G,
                             const int32_t* unit_col, const double* alpha, const int32_t* hist,
                             const int32_t* hist_cnt, double* FR, double* HDR, double* FR2, int model,
                             const double* mp0, const double* mp1) {
    if (KP == 12 && n > 512 && n <= PFB_K2_RPT * PFB_K2_THREADS_REG && FR2 != nullptr && pfb_k2_regcols_fits(KP, n)) {
        // two CTAs per SM: panel without the two extra columns (registers), 256 threads
        auto kern = pfb_k2_woodbury_build<KP, true, true>;
        const size_t smem = (size_t)KP * (size_t)((n + 1) | 1) * 8 + (size_t)PFB_K2_SCRATCH_REG(KP) * 8;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<U, PFB_K2_THREADS_REG, smem, st>>>(n, J, X, G, unit_col, alpha, hist, hist_cnt, nullptr, HDR, FR2, model, mp0,
                                                  mp1, u_base);
        return cudaGetLastError();
    }
    if (pfb_k2_uses_smem_panel(KP, n)) {
        int threads = n >= PFB_K2_THREADS_SMEM ? PFB_K2_THREADS_SMEM : ((n + 31) / 32) * 32;
        if (threads < 64) threads = 64;  // >= KP threads are needed by the small-matrix phases
        auto kern = pfb_k2_woodbury_build<KP, true>;
        const size_t smem = k2_panel_bytes(KP, n) + k2_scratch_bytes(KP);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<U, threads, smem, st>>>(n, J, X, G, unit_col, alpha, hist, hist_cnt, nullptr, HDR, FR2, model, mp0, mp1,
                                       u_base);
        return cudaGetLastError();
    }
    if (FR == nullptr) return cudaErrorInvalidValue;
    int threads = n >= PFB_K2_THREADS ? PFB_K2_THREADS : ((n + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    // static arrays + scratch exceed the 48 KB default at KP = 24: opt in like the shared-panel variant
    auto kern = pfb_k2_woodbury_build<KP, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2_scratch_bytes(KP));
    if (e != cudaSuccess) return e;
    kern<<<U, threads, k2_scratch_bytes(KP), st>>>(n, J, X, G, unit_col, alpha, hist, hist_cnt, FR, HDR, FR2, model, mp0,
                                                   mp1, u_base);
    return cudaGetLastError();
}

// units [u_base, u_base + U)
extern "C" cudaError_t pfb_launch_k2_range(cudaStream_t st, int KP, int n, int u_base, int U, int J, const double* X,
                                           const double* G, const int32_t* unit_col, const double* alpha,
                                           const int32_t* hist, const int32_t* hist_cnt, double* FR,
                                           double* HDR, double* FR2, int model, const double* mp0,
                                           const double* mp1);
extern "C" cudaError_t pfb_launch_k2(cudaStream_t st, int KP, int n, int U, int J, const double* X,
                                     const double* G, const int32_t* unit_col, const double* alpha,
                                     const int32_t* hist, const int32_t* hist_cnt, double* FR,
                                     double* HDR, double* FR2, int model, const double* mp0,
                                     const double* mp1) {
    return pfb_launch_k2_range(st, KP, n, 0, U, J, X, G, unit_col, alpha, hist, hist_cnt, FR, HDR, FR2, model, mp0,
                               mp1);
}
extern "C" cudaError_t pfb_launch_k2_range(cudaStream_t st, int KP, int n, int u_base, int U, int J, const double* X,
                                           const double* G, const int32_t* unit_col, const double* alpha,
                                           const int32_t* hist, const int32_t* hist_cnt, double* FR,
                                           double* HDR, double* FR2, int model, const double* mp0,
                                           const double* mp1) {
    if (U <= 0) return cudaSuccess;
    switch (KP) {
#define PFB_K2_CASE(k) \
    case k:            \
        return launch_k2<k>(st, n, u_base, U, J, X, G, unit_col, alpha, hist, hist_cnt, FR, HDR, FR2, model, mp0, mp1);
        PFB_K2_CASE(12)
        PFB_K2_CASE(20)
        PFB_K2_CASE(24)
#undef PFB_K2_CASE
    }
    return cudaErrorInvalidValue;
}
