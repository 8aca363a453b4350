// K3 elbo_sample_fused on the FP64 tensor cores (and K5 materialize_best: the same kernel with a
// draws pointer).
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
// Mapping.  A warp owns 8 * DS draws; lane (g = lane / 4, t = lane % 4) owns, for draw g of each
// draw set, the rows {8 b + 2 t, 8 b + 2 t + 1} of every 8-row block b — exactly ONE Philox4x32-7
// call (row pair 4 b + t, draw pair {k, k + 8}: pf_rng.h) per block for the lane's four variates of
// its two draw sets.  With that ownership both halves of the Q-apply are
// DMMA m8n8k4 products whose A / C fragments are the lane's own normals:
//   pass 0   w[draw][j]  += sum_rows  u~[draw][row] * Vh[row][j]      A = u~ (8 draws x 4 rows),
//                                                                     B = Vh  (4 rows x 8 j)
//   pass 1   z[draw][row] = u~[draw][row] - sum_j c[draw][j] Vh[row][j]   C = u~, A = -c = -T w,
//                                                                     B = Vh' (4 j x 8 rows)
// then x = sqrt(alpha) .* z + mu and the model's per-element accumulation.  The normals are
// regenerated in pass 1 from the counter-based RNG instead of being stored.
//
// The unit's factor record (rows of RS2 doubles {Vh[0..KP), sqrt(alpha), mu, pad}, 4-double groups
// XOR-swizzled per row so both fragment access patterns are bank-conflict free; written by K2) is
// brought into shared memory by 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx) in chunks of
// PFB_K3_RC rows: all chunks stay resident when they fit (n <= ~1150 at history 6), otherwise they
// stream through a ring, twice per sweep.  A CTA loops over the sweeps (256 draws) of its unit so
// the record is loaded once per unit.  The 1024-layer ziggurat table (8-byte packed entries, one
// LDS.64 per variate) sits in shared memory four times, interleaved entry by entry (copy = lane % 4),
// which spreads the random layer lookups of a half-warp over the bank pairs.  Elements that leave
// the ziggurat fast path (0.43 %) keep their PROVISIONAL table value in the tensor-core sums and are
// queued; the queue is finished 32 elements at a time (one element per lane: the real ziggurat
// continuation) and only the differences (true - provisional; zero for the half that the wedge test
// accepts) are folded in, column-parallel, through the warp's shared correction table.  With a
// resident record the queue is flushed once per sweep.
// Lean mode writes 16 B per draw (logp, logq); with a draws pointer x is written as well.
#include "pfb_common.cuh"
#include "pf_rng.h"

#ifndef PFB_K3_KP
#define PFB_K3_KP 12
#define PFB_K3_ENTRY pfb_launch_k3_kp12
#endif

#ifndef PFB_K3_SWP
#define PFB_K3_SWP 0  // software-pipeline the normals of block o+1 under the DMMAs of block o
#endif
#ifndef PFB_K3_STAGGER
#define PFB_K3_STAGGER 0  // ns of start delay for every other warp of a scheduler
#endif
#define PFB_K3_MAXWARPS 16
#define PFB_K3_DS 2      // draw sets (8 draws each) per warp
#define PFB_K3_RC 128    // record rows per TMA chunk (16 blocks of 8 rows)
#define PFB_K3_DCAP 32   // deferred-list capacity per warp and round
#define PFB_K3_ZREP 2    // interleaved copies of the 1024-layer ziggurat table (8-byte entries) in shared memory (16 KB)
#define PFB_K3_REQCAP 128  // queued slow-path elements per warp (a sweep of n = 1024 queues ~70)
static_assert(PF_ZIG_LAYERS == 1024 && PFB_K3_ZREP == 2, "pfb_zig_fast_rep address arithmetic");

__device__ __forceinline__ double pfb_lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 pfb_lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// acc += x * x unless flag != 0 (one predicated DFMA)
__device__ __forceinline__ void pfb_sqacc_unless(double& acc, double x, uint32_t flag) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t@p fma.rn.f64 %0, %1, %1, %0;\n\t}" : "+d"(acc) : "d"(x), "r"(flag));
}

// Column selection (MODE 1 only): instead of all K draws of a slot's unit, produce the draws
// list[slot * cap + j].x (j < cnt[slot]) and write each to column list[..].y of draws_out — the
// resampled draws regenerated on demand, so that the whole pool need not be materialised.
struct pfb_k3_sel {
    const int32_t* cnt;  // [nslots] or nullptr = all K draws
    const int2* list;    // [nslots x cap] (draw index, output column)
    int cap;
};

// Paths whose optimisation recorded no iteration (L = 0): the reference still draws from
// fit_distributions[1] = N(theta_0 + H_0 grad_0, H_0), H_0 = I (src/singlepath.jl:224-228 with
// fit_iteration = 0; src/inverse_hessian.jl:38-40).  Slots with unit < 0 draw from that normal when
// this block is given, and write NaN otherwise.
struct pfb_k3_fb {
    const double* X;           // trajectory points / gradients, n x T column-major
    const double* G;
    const int64_t* point_off;  // [P + 1] first column of every path
    const uint64_t* seeds;     // [P] one seed per path
    const int32_t* path_of_slot;  // nullptr: slot == path
};

struct pfb_model_params {
    const double* p0;  // DIAGNORMAL: mean[n]
    const double* p1;  // DIAGNORMAL: 1/sd[n]
    double c0;         // DIAGNORMAL: -sum(log sd) - n/2 log(2 pi)
};

template <int MODEL>
struct pfb_model_acc {
    double a, b;
    __device__ __forceinline__ void init() { a = 0.0; b = 0.0; }
    // general element (any row index i < n)
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
    // element with i > 0 guaranteed (body blocks)
    __device__ __forceinline__ void add_nz(int i, double x, const pfb_model_params& mp) {
        if (MODEL == PFB_MODEL_DIAGNORMAL) add(i, x, mp); else a = fma(x, x, a);
    }
    __device__ __forceinline__ void add_nz_unless(int i, double x, const pfb_model_params& mp, uint32_t flag) {
        if (MODEL == PFB_MODEL_DIAGNORMAL) {
            if (!flag) add(i, x, mp);
        } else {
            pfb_sqacc_unless(a, x, flag);
        }
    }
    // combine the partial accumulators of the 4 lanes that share a draw
    __device__ __forceinline__ void group_reduce() {
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        b += __shfl_xor_sync(0xffffffffu, b, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        b += __shfl_xor_sync(0xffffffffu, b, 2);
    }
    __device__ __forceinline__ double finish(int n, const pfb_model_params& mp) const {
        if (MODEL == PFB_MODEL_EXTERNAL) return 0.0;  // log p comes from K8 (GEMM over the draws)
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

struct pfb_k3_warp_list {
    // pass 0: the queue of elements that left the ziggurat fast path, and one batch of finished ones
    uint32_t req[PFB_K3_REQCAP];  // row | owner lane << 20 | draw set << 25
    double dz[33];                // true - provisional variate (33: the four arrays start in different banks)
    double pdz[33];               // p_row * dz                     (single-pass statistics)
    double dq[33];                // true^2 - provisional^2
    double cq[33];                // p_row * dq + 2 r_row * dz      (single-pass statistics)
    uint32_t bmeta[32];           // the batch's req words (0xFFFFFFFF: nothing to fold)
};
// pass 1 (two-pass modes): per-chunk list of finished variates; shares the batch arrays' memory
struct pfb_k3_warp_list1 {
    uint32_t req_unused[PFB_K3_REQCAP];
    double z[PFB_K3_DCAP];
    uint32_t meta[PFB_K3_DCAP];   // row offset inside the chunk | owner lane << 16 | draw set << 24
};
static_assert(sizeof(pfb_k3_warp_list1) <= sizeof(pfb_k3_warp_list), "pass-1 list aliases the pass-0 batch");

__device__ __forceinline__ void pfb_dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

// Fast ziggurat step on the replicated table (zig_base = shared-memory byte address of this
// lane's table copy; 32-bit word layout and the packed 8-byte entry: pf_rng.h).  Returns 0 and the
// variate in z when accepted, 1 and the PROVISIONAL table value in z when the element has to take the
// slow path (the caller queues it; its outputs are corrected or overwritten later).
__device__ __forceinline__ uint32_t pfb_zig_fast_rep(uint32_t w, uint32_t zig_base, double& z) {
    uint32_t elo, ehi;
    const uint32_t addr = ((w >> 17) & 0x3FF0u) + zig_base;  // layer (bits 21-30) * 16 = 2 copies * 8 B
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(elo), "=r"(ehi) : "r"(addr));
    uint32_t mh, kqh;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(mh) : "r"(w), "r"(0xFFFFFu), "r"(0x3FF00000u));    // (a&b)|c
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(kqh) : "r"(elo), "r"(0xFFFFFu), "r"(0x3FF00000u));
    const double m = __hiloint2double((int)mh, 0);          // 1 + j 2^-20
    const double xe = __hiloint2double((int)ehi, (int)elo);
    const double x = fma(m, xe, -xe);                       // j 2^-20 x_i, rounded once
    z = __hiloint2double(__double2hiint(x) ^ (int)(w & 0x80000000u), __double2loint(x));
    return mh >= kqh ? 1u : 0u;
}

template <int V>
struct pfb_ic {
    static constexpr int value = V;
};

// MODE 0: two passes, lean (logp, logq only)      MODE 1: two passes, x materialised (K5 / mode M)
// MODE 2: single pass for the diagonal-quadratic target family (pfb_common.cuh, PFB_HDR_M): the
//         normals are generated ONCE; S = sum d (x - m)^2 and x_0 follow from linear functionals of
//         u~ (w = Vh' u~, Vh' (p u~), r' u~ — all DMMA columns) and sum p u~^2, so the second
//         Philox/ziggurat sweep (the kernel's bottleneck) disappears.
template <int KP, int MODEL, int MODE, bool SELM = false>
__global__ void __launch_bounds__(PFB_K3_MAXWARPS * 32, 1)
pfb_k3_elbo_sample(int n, int K, int splits, int NS, const int32_t* __restrict__ unit_list,
                   const double* __restrict__ FR2, const double* __restrict__ HDR,
                   const uint64_t* __restrict__ seeds, const double* __restrict__ u_host,
                   pfb_model_params mp, double* __restrict__ logp_out, double* __restrict__ logq_out,
                   double* __restrict__ draws_out, pfb_k3_sel sel, pfb_k3_fb fb) {
    constexpr bool MATERIALIZE = (MODE == 1);
    constexpr bool QUAD = (MODE == 2);
    constexpr int RS2 = (KP == 12) ? 16 : 32;
    constexpr int RC = PFB_K3_RC;
    constexpr int DS = PFB_K3_DS;
    constexpr int NT0 = QUAD ? (2 * KP + 7) / 8 : (KP + 7) / 8;  // column tiles of pass 0
    constexpr int NS1 = KP / 4;        // j steps of pass 1
    constexpr int HB = (KP + 7) / 8;   // head blocks (rows < KP get the Vc' multiply)
    constexpr int NHP = (KP / 2 + 3) / 4;  // head row pairs generated per lane
    constexpr int CWW = QUAD ? 2 * KP : KP;  // per-draw shared accumulators: w (and Vh'(p u~))
    constexpr int CWS = CWW + (QUAD ? 2 : 1);  // + the slow-path corrections of |u|^2 (and of the quadratic statistic)
    static_assert(DS == 2, "pending nibbles assume two draw sets");
    static_assert(RC / 8 * 4 <= 64, "pending mask is 64 bits per chunk");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sStage = reinterpret_cast<double*>(smem_raw);                      // NS * RC * RS2
    uint64_t* sZig = reinterpret_cast<uint64_t*>(sStage + (size_t)NS * RC * RS2);  // 1024 * 2, copy-interleaved
    double* sT = reinterpret_cast<double*>(sZig + PF_ZIG_LAYERS * PFB_K3_ZREP);  // KP*KP
    double* sVc = sT + KP * KP;                                                // KP*KP
    double* sM = sVc + KP * KP;                                                // KP*KP + KP (QUAD)
    double* sC = sM + (QUAD ? KP * KP + KP : 0);                               // NW * DS * 8 * CWW
    pfb_k3_warp_list* sList = reinterpret_cast<pfb_k3_warp_list*>(sC + (size_t)NW * DS * 8 * CWS);
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sList + NW);                  // NS

    const int slot = blockIdx.x / splits;
    const int split = blockIdx.x - slot * splits;
    const int unit = unit_list ? unit_list[slot] : slot;
    const int DPS = NW * DS * 8;                 // draws per sweep
    constexpr bool SEL = SELM;  // column selection is its own instantiation: the other kernels are untouched
    const int2* sel_list = SEL ? sel.list + (int64_t)slot * sel.cap : nullptr;
    const int Kslot = SEL ? sel.cnt[slot] : K;   // draws this slot produces
    const int S = (Kslot + DPS - 1) / DPS;       // sweeps of this unit
    if (SEL && split >= S) return;               // (column selection: most CTAs of a slot have nothing to do)
    if (unit < 0 && fb.X != nullptr) {
        // identity fit of iteration 0: x = theta_0 + grad_0 + u, log q = -(n log 2 pi + |u|^2) / 2; one warp per draw
        const int path = fb.path_of_slot ? fb.path_of_slot[slot] : slot;
        const double* x0 = fb.X + fb.point_off[path] * (int64_t)n;
        const double* g0 = fb.G + fb.point_off[path] * (int64_t)n;
        const uint64_t fseed = fb.seeds[path];
        const uint32_t f0 = (uint32_t)fseed, f1 = (uint32_t)(fseed >> 32);
        for (int sw = split; sw < S; sw += splits) {
            const int ka = sw * DPS, kb = min(Kslot, (sw + 1) * DPS);
            for (int j = ka + warp; j < kb; j += NW) {
                const uint32_t kdraw = SEL ? (uint32_t)sel_list[j].x : (uint32_t)j;
                const int64_t oc = SEL ? (int64_t)sel_list[j].y : (int64_t)slot * K + j;
                double usq = 0.0;
                pfb_model_acc<MODEL> ma;
                ma.init();
                for (int rp = lane; 2 * rp < n; rp += 32) {
                    double z0, z1;
                    pf_normal_pair((uint32_t)rp, kdraw, f0, f1, PF_ZIG_XK_DEV, PF_ZIG_F_DEV, &z0, &z1);
                    const int i = 2 * rp;
                    const double xa = (x0[i] + g0[i]) + z0;
                    usq = fma(z0, z0, usq);
                    ma.add(i, xa, mp);
                    if (MATERIALIZE && draws_out) draws_out[oc * n + i] = xa;
                    if (i + 1 < n) {
                        const double xb = (x0[i + 1] + g0[i + 1]) + z1;
                        usq = fma(z1, z1, usq);
                        ma.add(i + 1, xb, mp);
                        if (MATERIALIZE && draws_out) draws_out[oc * n + i + 1] = xb;
                    }
                }
                usq = pfb_warp_sum(usq);
                ma.a = pfb_warp_sum(ma.a);
                ma.b = pfb_warp_sum(ma.b);
                if (lane == 0 && !SEL) {
                    if (logp_out) logp_out[oc] = ma.finish(n, mp);
                    if (logq_out) logq_out[oc] = (fma((double)n, PFB_LOG2PI, 0.0) + usq) / -2.0;
                }
            }
        }
        return;
    }
    if (unit < 0) {  // path without a usable iteration (K5 only): no fitted normal, NaN draws
        for (int sw = split; sw < S; sw += splits) {
            const int ka = sw * DPS, kb = min(Kslot, (sw + 1) * DPS);
            if (!SEL) {
                for (int k = ka + tid; k < kb; k += blockDim.x) {
                    if (logp_out) logp_out[(int64_t)slot * K + k] = NAN;
                    if (logq_out) logq_out[(int64_t)slot * K + k] = NAN;
                }
                if (MATERIALIZE && draws_out != nullptr && kb > ka) {
                    double* d0 = draws_out + ((int64_t)slot * K + ka) * n;
                    for (int64_t e = tid; e < (int64_t)(kb - ka) * n; e += blockDim.x) d0[e] = NAN;
                }
            } else {
                for (int j = ka; j < kb; ++j) {
                    double* d0 = draws_out + (int64_t)sel_list[j].y * n;
                    for (int e = tid; e < n; e += blockDim.x) d0[e] = NAN;
                }
            }
        }
        return;
    }
    const int npad = pfb_npad8(n);
    const double* fr = FR2 + (int64_t)unit * npad * RS2;
    const double* hdr = HDR + (int64_t)unit * pfb_hs_of(KP);
    const uint64_t seed = seeds[unit];
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const bool HOST_U = (u_host != nullptr);

    const int C = (npad + RC - 1) / RC;  // chunks per pass
    const bool resident = (C <= NS);
    const int nsweeps_cta = (S - split + splits - 1) / splits;
    const int Q = resident ? C : (QUAD ? 1 : 2) * C * nsweeps_cta;  // TMA loads issued by this CTA

    for (int e = tid; e < KP * KP; e += blockDim.x) {
        sT[e] = hdr[e];
        sVc[e] = hdr[KP * KP + e];
    }
    if (QUAD)
        for (int e = tid; e < KP * KP + KP; e += blockDim.x) sM[e] = hdr[PFB_HDR_M(KP) + e];  // M, then rv
    for (int e = tid; e < PF_ZIG_LAYERS * PFB_K3_ZREP; e += blockDim.x) sZig[e] = PF_ZIG_XK_DEV[e / PFB_K3_ZREP];
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) pfb_mbar_init(&sBar[s], 1);
        pfb_fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int q) {
        const int c = q % C, s = q % NS;
        const int r0 = c * RC;
        const int rows = min(RC, npad - r0);
        const uint32_t bytes = (uint32_t)(rows * RS2 * 8);
        pfb_mbar_expect_tx(&sBar[s], bytes);
        pfb_tma_load_1d(sStage + (size_t)s * RC * RS2, fr + (int64_t)r0 * RS2, bytes, &sBar[s]);
    };
    if (tid == 0) {
        for (int q = 0; q < NS && q < Q; ++q) issue(q);
    }
    const double logdet = hdr[PFB_HDR_LOGDET(KP)];
    const bool pd_ok = hdr[PFB_HDR_FLAG(KP)] != 0.0;
    const int H = min(KP, n);          // head rows
    const bool tail_special = (n & 7) != 0;

    // lane-constant fragment offsets (bytes) inside an 8-row block.
    // pass 0: column cc = 8h + g of row 2t+e;  two-pass: cc = j (Vh[row][j]);  single pass: cc < KP
    // is Vh[row][cc], cc >= KP is p_row * Vh[row][cc - KP].
    uint32_t off0[2][NT0];
    bool scaled[NT0];  // this lane's column of tile h is a p-scaled one (QUAD)
#pragma unroll
    for (int h = 0; h < NT0; ++h) {
        const int cc = 8 * h + g;
        scaled[h] = QUAD && cc >= KP;
        const int j = scaled[h] ? cc - KP : cc;
#pragma unroll
        for (int e = 0; e < 2; ++e)
            off0[e][h] = 8u * (uint32_t)((2 * t + e) * RS2 + ((j < RS2 ? j : 0) ^ pfb_swz(2 * t + e)));
    }
    uint32_t off1[NS1];     // pass 1: Vh[row g][j = 4s+t]
#pragma unroll
    for (int s = 0; s < NS1; ++s) off1[s] = 8u * (uint32_t)(g * RS2 + ((4 * s + t) ^ pfb_swz(g)));
    uint32_t offam[2];      // {sqrt(alpha), mu} of row 2t+e;  {p, r} sit 16 bytes further
#pragma unroll
    for (int e = 0; e < 2; ++e) offam[e] = 8u * (uint32_t)((2 * t + e) * RS2 + (KP ^ pfb_swz(2 * t + e)));

    const uint32_t zig_base = pfb_smem_u32(sZig + (lane & (PFB_K3_ZREP - 1)));  // copy = lane % 2: 8-byte stride
    pfb_k3_warp_list& wl = sList[warp];
    pfb_k3_warp_list1& wl1 = *reinterpret_cast<pfb_k3_warp_list1*>(&sList[warp]);
    double* cw = sC + (size_t)warp * DS * 8 * CWS;  // this warp's [DS][8][CWS]: w (| Vh'(p u~)), then c = T w

    int q = 0;  // TMA load sequence number of the next chunk to consume (ring mode)
    bool first_pass = true;
#pragma unroll 1
    for (int sw = split; sw < S; sw += splits) {
        uint32_t kd[DS];      // draw index (clamped)
        int64_t ocol[DS];     // MODE 1: output column of the draw (the lean modes recompute it at the end)
        uint32_t actmask = 0u;  // pending-nibble mask of the active draw sets (bit d*2+e)
        // the Philox counter of the lane's two draws (k, k + 8) in an all-draws sweep (pf_rng.h)
        const uint32_t dpair = pf_draw_pair((uint32_t)(sw * DPS + warp * DS * 8 + g));
#pragma unroll
        for (int d = 0; d < DS; ++d) {
            const int kraw = sw * DPS + (warp * DS + d) * 8 + g;
            if (kraw < Kslot) actmask |= 3u << (2 * d);
            const int pos = kraw < Kslot ? kraw : Kslot - 1;
            kd[d] = SEL ? (uint32_t)sel_list[pos].x : (uint32_t)pos;
            if (MATERIALIZE) ocol[d] = SEL ? (int64_t)sel_list[pos].y : (int64_t)slot * K + kd[d];
        }
        double unormsq[DS];
        pfb_model_acc<MODEL> macc[DS];
        double wacc[DS][NT0][2];  // pass 0 accumulators: column cc = 8h + 2t + {0,1} of draw g
        double ncf[DS][NS1];      // pass 1 A fragments: -c[draw g][j = 4s + t]
        double qsum[DS], u0[DS];  // QUAD: sum u~ (p u~ + 2 r) = sum p u~^2 + 2 sum r u~;  u~_0 (lane t = 0)
#pragma unroll
        for (int d = 0; d < DS; ++d) {
            unormsq[d] = 0.0;
            qsum[d] = u0[d] = 0.0;
            macc[d].init();
#pragma unroll
            for (int h = 0; h < NT0; ++h) wacc[d][h][0] = wacc[d][h][1] = 0.0;
#pragma unroll
            for (int s1 = 0; s1 < NS1; ++s1) ncf[d][s1] = 0.0;
        }
        for (int e = lane; e < DS * 8 * CWS; e += 32) cw[e] = 0.0;  // slow-path corrections of w, |u|^2, ...
        int nreq = 0;  // queued slow-path elements of this warp (pass 0)
        __syncwarp();

        // Pass 0: finish the queued elements, 32 at a time (one per lane: the real ziggurat continuation),
        // and fold the DIFFERENCES to their provisional values into the warp's correction table:
        // lane c owns column c (w, Vh'(p u~), then the |u|^2 and quadratic-statistic corrections) and
        // walks the batch in queue order (fixed order => deterministic).  rows: record rows in shared
        // memory, row0: the record row of rows[0].
        auto flush_queue = [&](const double* rows, int row0) {
            __syncwarp();
#pragma unroll 1
            for (int b0 = 0; b0 < nreq; b0 += 32) {
                const int it = b0 + lane;
                if (it < nreq) {
                    const uint32_t mt = wl.req[it];
                    const uint32_t row = mt & 0xFFFFFu;
                    const int kraw = sw * DPS + (warp * DS + (int)(mt >> 25)) * 8 + (int)((mt >> 22) & 7u);
                    const uint32_t kdraw = SEL ? (uint32_t)sel_list[kraw].x : (uint32_t)kraw;
                    const pf_slow_t fz = pf_normal_finish_slow(row, kdraw, k0, k1, PF_ZIG_XK_DEV, PF_ZIG_F_DEV);
                    const double dz = fz.z - fz.zprov;
                    const double dq = dz * (fz.z + fz.zprov);
                    wl.dz[lane] = dz;
                    wl.dq[lane] = dq;
                    if (QUAD) {
                        const double* rr = rows + ((int)row - row0) * RS2;
                        const int swz = pfb_swz((int)row);
                        const double pp = rr[(KP + 2) ^ swz];
                        wl.pdz[lane] = pp * dz;
                        wl.cq[lane] = fma(pp, dq, rr[(KP + 3) ^ swz] * dz);  // slot KP + 3 holds 2 r
                    }
                    wl.bmeta[lane] = dz != 0.0 ? mt : 0xFFFFFFFFu;  // the wedge test accepted it: nothing changes
                }
                __syncwarp();
                const int nbt = min(32, nreq - b0);
#pragma unroll
                for (int cc = lane; cc < CWS; cc += 32) {  // CWS <= 50: at most two columns per lane
                    const int col = cc < KP ? cc : cc - KP;
                    const double* val = cc < KP ? wl.dz : (cc < CWW ? wl.pdz : (cc == CWW ? wl.dq : wl.cq));
#pragma unroll 1
                    for (int j = 0; j < nbt; ++j) {
                        const uint32_t mt = wl.bmeta[j];
                        if (mt == 0xFFFFFFFFu) continue;
                        const int row = (int)(mt & 0xFFFFFu);
                        const double v = cc < CWW ? rows[(row - row0) * RS2 + (col ^ pfb_swz(row))] : 1.0;
                        double* cv = cw + ((mt >> 25) * 8 + ((mt >> 22) & 7u)) * CWS + cc;
                        *cv = fma(v, val[j], *cv);
                    }
                }
                __syncwarp();
            }
        };

        // One pass over all chunks.  PS::value = 0: accumulate w = Vh' u~ (QUAD: and the other
        // statistics);  1: x and the model sums.
        auto run_pass = [&](auto PS) {
            constexpr int PASS = decltype(PS)::value;
            // ---- head: u~[0..H) = Vc' u[0..H)  (src/woodbury.jl:139); recomputed in each pass ------
            double zhd[DS][HB][2];  // u~ of this lane's head rows
#pragma unroll
            for (int d = 0; d < DS; ++d) {
                double own[NHP][2];
#pragma unroll
                for (int r = 0; r < NHP; ++r) {
                    const int jj = 4 * r + t;  // row pair generated by this lane
                    double z0 = 0.0, z1 = 0.0;
                    if (2 * jj < H) {
                        if (HOST_U) {
                            const double* uh = u_host + ((int64_t)unit * K + kd[d]) * n;
                            z0 = uh[2 * jj];
                            if (2 * jj + 1 < H) z1 = uh[2 * jj + 1];
                        } else {
                            pf_normal_pair((uint32_t)jj, kd[d], k0, k1, PF_ZIG_XK_DEV, PF_ZIG_F_DEV, &z0, &z1);
                            if (2 * jj + 1 >= H) z1 = 0.0;
                        }
                    }
                    own[r][0] = z0;
                    own[r][1] = z1;
                    if (PASS == 0) {
                        unormsq[d] = fma(z0, z0, unormsq[d]);
                        unormsq[d] = fma(z1, z1, unormsq[d]);
                    }
                }
                double zh[KP];
#pragma unroll
                for (int jj = 0; jj < KP / 2; ++jj) {
                    const int srcl = (lane & ~3) | (jj & 3);
                    zh[2 * jj] = __shfl_sync(0xffffffffu, own[jj >> 2][0], srcl);
                    zh[2 * jj + 1] = __shfl_sync(0xffffffffu, own[jj >> 2][1], srcl);
                }
#pragma unroll
                for (int bb = 0; bb < HB; ++bb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 8 * bb + 2 * t + e;
                        double acc = 0.0;
                        if (j < KP) {
#pragma unroll
                            for (int m = 0; m < KP; ++m)
                                if (m <= j) acc = fma(sVc[m * KP + j], zh[m], acc);
                        }
                        zhd[d][bb][e] = acc;
                    }
                if (QUAD) u0[d] = zhd[d][0][0];  // u~ of row 0 (meaningful in lane t = 0)
            }

#pragma unroll 1
            for (int c = 0; c < C; ++c) {
                int s;
                if (resident) {
                    s = c;
                    if (first_pass) pfb_mbar_wait(&sBar[s], 0u);  // loaded once, stays resident
                } else {
                    s = q % NS;
                    pfb_mbar_wait(&sBar[s], (uint32_t)((q / NS) & 1));
                }
#if PFB_K3_STAGGER
                // warps of one scheduler alternate between an integer phase (Philox / ziggurat) and an
                // FP64 phase (DMMA); left alone they run in lock-step and the two pipes take turns
                // idling.  Delaying every other warp of a scheduler by half a block period locks the
                // pairs in anti-phase.
                if (c == 0 && ((warp >> 2) & 1)) __nanosleep(PFB_K3_STAGGER);
#endif
                const double* st = sStage + (size_t)s * RC * RS2;
                const uint32_t st_u32 = pfb_smem_u32(st);
                const int r0 = c * RC;
                const int nb = (min(npad, r0 + RC) - r0) >> 3;
                // pending (deferred) elements: nibble per block, bit d*2+e; the first block of the
                // chunk ends up in the highest of the nb nibbles
                unsigned long long pend = 0ull;

                // normals of a body block (fast ziggurat step; rejected elements: z = 0, nibble bit)
                auto gen = [&](int o, double (&z)[DS][2], uint32_t& nib) {
                    const int b = (r0 >> 3) + o;
                    nib = 0u;
                    if (!SEL) {
                        // ONE Philox4x32-7 call: the lane's 2 rows x 2 draw sets (draws k, k + 8)
                        uint32_t w4[4];
                        pf_bits4((uint32_t)(4 * b + t), 0u, dpair, k0, k1, 0u, w4);
#pragma unroll
                        for (int d = 0; d < DS; ++d) {
                            nib |= pfb_zig_fast_rep(w4[2 * d], zig_base, z[d][0]) << (2 * d);
                            nib |= pfb_zig_fast_rep(w4[2 * d + 1], zig_base, z[d][1]) << (2 * d + 1);
                        }
                    } else {
#pragma unroll
                        for (int d = 0; d < DS; ++d) {  // arbitrary draws: one call per draw set
                            uint32_t w4[4];
                            pf_bits4((uint32_t)(4 * b + t), 0u, pf_draw_pair(kd[d]), k0, k1, 0u, w4);
                            const bool hi = pf_draw_half(kd[d]) != 0u;
                            nib |= pfb_zig_fast_rep(hi ? w4[2] : w4[0], zig_base, z[d][0]) << (2 * d);
                            nib |= pfb_zig_fast_rep(hi ? w4[3] : w4[1], zig_base, z[d][1]) << (2 * d + 1);
                        }
                    }
                };
                // SP = 0: consume the normals zin/nibin generated by gen();  SP = 1: special block
                auto block = [&](auto SP, int o, const double (&zin)[DS][2], uint32_t nibin) {
                    constexpr bool SPECIAL = decltype(SP)::value != 0;
                    const int b = (r0 >> 3) + o;
                    const uint32_t blk = st_u32 + (uint32_t)(o * 8 * RS2 * 8);
                    const int R = r0 + o * 8 + 2 * t;  // this lane's first row of the block
                    double z[DS][2];
                    uint32_t nib = 0u;
                    if (!SPECIAL) {
#pragma unroll
                        for (int d = 0; d < DS; ++d) {
                            z[d][0] = zin[d][0];
                            z[d][1] = zin[d][1];
                        }
                        nib = nibin;
                    } else {
#pragma unroll
                        for (int d = 0; d < DS; ++d) {
                            z[d][0] = z[d][1] = 0.0;
                            if (R < H) {
                                // head rows come in pairs (H is a multiple of 4 unless n < KP, and
                                // then rows >= n are zero)
#pragma unroll
                                for (int bb = 0; bb < HB; ++bb)
                                    if (bb == b) {
                                        z[d][0] = zhd[d][bb][0];
                                        z[d][1] = (R + 1 < H) ? zhd[d][bb][1] : 0.0;
                                    }
                            } else if (R < n) {
                                if (HOST_U) {
                                    const double* uh = u_host + ((int64_t)unit * K + kd[d]) * n;
                                    z[d][0] = uh[R];
                                    if (R + 1 < n) z[d][1] = uh[R + 1];
                                } else {
                                    uint32_t w4[4];
                                    pf_bits4((uint32_t)(4 * b + t), 0u, pf_draw_pair(kd[d]), k0, k1, 0u, w4);
                                    const bool hi = pf_draw_half(kd[d]) != 0u;
                                    nib |= pfb_zig_fast_rep(hi ? w4[2] : w4[0], zig_base, z[d][0]) << (2 * d);
                                    if (R + 1 < n) {
                                        nib |= pfb_zig_fast_rep(hi ? w4[3] : w4[1], zig_base, z[d][1]) << (2 * d + 1);
                                    } else {
                                        z[d][1] = 0.0;
                                    }
                                }
                            }
                        }
                    }
                    nib &= actmask;  // (inactive draws are never queued; their sums are discarded)
                    pend = (pend << 4) | nib;
                    if (PASS == 0) {
                        const bool headrow = SPECIAL && R < H;  // |u|^2 of head rows: added with the raw normals
                        double2 pr[2];
                        if (QUAD) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) pr[e] = pfb_lds128(blk + offam[e] + 16u);  // {p, r}
                        }
#pragma unroll
                        for (int d = 0; d < DS; ++d)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                if (QUAD) {
                                    // record slot KP + 3 holds 2 r: one fused chain per statistic
                                    if (!headrow) unormsq[d] = fma(z[d][e], z[d][e], unormsq[d]);
                                    qsum[d] = fma(fma(pr[e].x, z[d][e], pr[e].y), z[d][e], qsum[d]);
                                } else if (!headrow) {
                                    unormsq[d] = fma(z[d][e], z[d][e], unormsq[d]);
                                }
                            }
                        double vf[2][NT0];
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int h = 0; h < NT0; ++h) {
                                vf[e][h] = pfb_lds64(blk + off0[e][h]);
                                if (QUAD && 8 * h + 7 >= KP)  // tile holds p-scaled columns
                                    vf[e][h] *= (8 * h >= KP || scaled[h]) ? pr[e].x : 1.0;
                            }
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int d = 0; d < DS; ++d)
#pragma unroll
                                for (int h = 0; h < NT0; ++h)
                                    pfb_dmma(wacc[d][h][0], wacc[d][h][1], z[d][e], vf[e][h]);
                    } else {
                        double bf[NS1];
#pragma unroll
                        for (int s1 = 0; s1 < NS1; ++s1) bf[s1] = pfb_lds64(blk + off1[s1]);
                        double2 am[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) am[e] = pfb_lds128(blk + offam[e]);
#pragma unroll
                        for (int d = 0; d < DS; ++d) {
                            double d0 = z[d][0], d1 = z[d][1];
#pragma unroll
                            for (int s1 = 0; s1 < NS1; ++s1) pfb_dmma(d0, d1, ncf[d][s1], bf[s1]);
                            const double x0 = fma(am[0].x, d0, am[0].y);
                            const double x1 = fma(am[1].x, d1, am[1].y);
                            const bool act = (actmask >> (2 * d)) & 1u;
                            double* xo = MATERIALIZE ? draws_out + ocol[d] * n : nullptr;
                            if (!SPECIAL) {
                                macc[d].add_nz_unless(R, x0, mp, nib & (1u << (2 * d)));
                                macc[d].add_nz_unless(R + 1, x1, mp, nib & (2u << (2 * d)));
                                if (MATERIALIZE && act) {
                                    if ((n & 1) == 0) {
                                        // both rows exist; a pending one is rewritten by the fix-up
                                        *reinterpret_cast<double2*>(xo + R) = make_double2(x0, x1);
                                    } else {
                                        xo[R] = x0;
                                        xo[R + 1] = x1;
                                    }
                                }
                            } else {
                                if (R < n) {
                                    if (!(nib & (1u << (2 * d)))) macc[d].add(R, x0, mp);
                                    if (MATERIALIZE && act) xo[R] = x0;
                                }
                                if (R + 1 < n) {
                                    if (!(nib & (2u << (2 * d)))) macc[d].add(R + 1, x1, mp);
                                    if (MATERIALIZE && act) xo[R + 1] = x1;
                                }
                            }
                        }
                    }
                };

                // leading special blocks (head rows, or everything in parity mode), fast body,
                // trailing special block (n not a multiple of 8)
                int o_fast = HOST_U ? nb : max(0, min(nb, HB - (r0 >> 3)));
                int o_end = nb;
                if (tail_special && r0 + nb * 8 == npad && o_end > o_fast) --o_end;
                int o = 0;
                double zc[DS][2] = {};
                uint32_t nibc = 0u;
#pragma unroll 1
                for (; o < o_fast; ++o) block(pfb_ic<1>{}, o, zc, 0u);
#if PFB_K3_SWP
                // software pipeline: the normals of block o+1 are generated in the same basic block
                // as the tensor-core work of block o
                if (o < o_end) {
                    double zn[DS][2];
                    uint32_t nibn;
                    gen(o, zc, nibc);
#pragma unroll 1
                    for (; o < o_end - 1; ++o) {
                        gen(o + 1, zn, nibn);
                        block(pfb_ic<0>{}, o, zc, nibc);
#pragma unroll
                        for (int d = 0; d < DS; ++d) {
                            zc[d][0] = zn[d][0];
                            zc[d][1] = zn[d][1];
                        }
                        nibc = nibn;
                    }
                    block(pfb_ic<0>{}, o, zc, nibc);
                    ++o;
                }
#else
#pragma unroll 1
                for (; o < o_end; ++o) {
                    gen(o, zc, nibc);
                    block(pfb_ic<0>{}, o, zc, nibc);
                }
#endif
#pragma unroll 1
                for (; o < nb; ++o) block(pfb_ic<1>{}, o, zc, 0u);

                // ---- elements that left the ziggurat fast path ------------------------------------------
                if (!HOST_U) {
                    const int cnt = __popcll(pend);
                    int incl = cnt;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        int v = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += v;
                    }
                    const int excl = incl - cnt;
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    // bit -> (block o, e, d): nibble index from the top, bit d*2+e inside
                    auto row_of = [&](int bit) { return (nb - 1 - (bit >> 2)) * 8 + 2 * t + (bit & 1); };
                    if (PASS == 0) {
                        // queue them (global row, owner lane, draw set); flushed when the queue is full,
                        // when the chunk is about to leave shared memory (ring mode), and after the last
                        // chunk of the pass (a resident record: the whole sweep's queue at once)
                        const bool chunk_flush = !resident || c == C - 1;
#pragma unroll 1
                        for (int done = 0;;) {
                            const int take = min(PFB_K3_REQCAP - nreq, total - done);
                            unsigned long long a0 = pend;
                            int flat = excl - done;
                            while (a0) {
                                const int bit = __ffsll((long long)a0) - 1;
                                a0 &= a0 - 1;
                                if (flat >= 0 && flat < take)
                                    wl.req[nreq + flat] = (uint32_t)(r0 + row_of(bit)) | ((uint32_t)lane << 20) |
                                                          ((uint32_t)((bit >> 1) & 1) << 25);
                                ++flat;
                            }
                            nreq += take;
                            done += take;
                            const bool last = done >= total;
                            if (nreq == PFB_K3_REQCAP || (last && chunk_flush && nreq > 0)) {
                                flush_queue(resident ? sStage : st, resident ? 0 : r0);
                                nreq = 0;
                            }
                            if (last) break;
                        }
                    } else {
                        // pass 1: the owner recomputes x of its queued elements from the finished variates
#pragma unroll 1
                        for (int base = 0; base < total; base += PFB_K3_DCAP) {
                            {
                                unsigned long long a0 = pend;
                                int flat = excl - base;
                                while (a0) {
                                    const int bit = __ffsll((long long)a0) - 1;
                                    a0 &= a0 - 1;
                                    if (flat >= 0 && flat < PFB_K3_DCAP)
                                        wl1.meta[flat] = (uint32_t)row_of(bit) | ((uint32_t)lane << 16) |
                                                        ((uint32_t)((bit >> 1) & 1) << 24);
                                    ++flat;
                                }
                            }
                            __syncwarp();
                            const int nitems = min(PFB_K3_DCAP, total - base);
                            for (int it = lane; it < nitems; it += 32) {
                                const uint32_t mt = wl1.meta[it];
                                const uint32_t row = (uint32_t)r0 + (mt & 0xFFFFu);
                                const int kraw = sw * DPS + (warp * DS + (int)(mt >> 24)) * 8 + (int)((mt >> 18) & 7u);
                                const uint32_t kdraw = SEL ? (uint32_t)sel_list[kraw].x : (uint32_t)kraw;
                                wl1.z[it] = pf_normal_finish_slow(row, kdraw, k0, k1, PF_ZIG_XK_DEV, PF_ZIG_F_DEV).z;
                            }
                            __syncwarp();
                            unsigned long long a0 = pend;
                            int flat = excl - base;
                            while (a0) {
                                const int bit = __ffsll((long long)a0) - 1;
                                a0 &= a0 - 1;
                                if (flat >= 0 && flat < PFB_K3_DCAP) {
                                    const int row = row_of(bit), dsi = (bit >> 1) & 1;
                                    const double* rr = st + row * RS2;
                                    const int swz = pfb_swz(row);
                                    const double* cv = cw + (dsi * 8 + g) * CWS;
                                    double zz = wl1.z[flat];
#pragma unroll
                                    for (int j = 0; j < KP; ++j) zz = fma(-rr[j ^ swz], cv[j], zz);
                                    const double2 am = *reinterpret_cast<const double2*>(rr + (KP ^ swz));
                                    const double x = fma(am.x, zz, am.y);
#pragma unroll
                                    for (int d = 0; d < DS; ++d)
                                        if (d == dsi) {
                                            macc[d].add(r0 + row, x, mp);
                                            if (MATERIALIZE)
                                                draws_out[ocol[d] * n + r0 + row] = x;
                                        }
                                }
                                ++flat;
                            }
                            __syncwarp();
                        }
                    }
                }
                if (!resident) {
                    __syncthreads();  // every warp is done with stage s
                    if (tid == 0 && q + NS < Q) issue(q + NS);
                    ++q;
                }
            }
        };

        run_pass(pfb_ic<0>{});
        first_pass = false;

        {
            // c = T w (upper triangular): the w fragments join the slow-path corrections in shared
            // memory so that every lane can form its pass-1 A fragments -c[4s + t]
            __syncwarp();
            if (t == 0) {
#pragma unroll
                for (int d = 0; d < DS; ++d) {
                    unormsq[d] += cw[(d * 8 + g) * CWS + CWW];
                    if (QUAD) qsum[d] += cw[(d * 8 + g) * CWS + CWW + 1];
                }
            }
#pragma unroll
            for (int d = 0; d < DS; ++d)
#pragma unroll
                for (int h = 0; h < NT0; ++h)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = 8 * h + 2 * t + e;
                        if (cc < CWW) cw[(d * 8 + g) * CWS + cc] += wacc[d][h][e];
                    }
            __syncwarp();
            double cf[DS][NS1];
#pragma unroll
            for (int d = 0; d < DS; ++d)
#pragma unroll
                for (int s1 = 0; s1 < NS1; ++s1) {
                    const int a = 4 * s1 + t;
                    double acc = 0.0;
#pragma unroll
                    for (int bcol = 0; bcol < KP; ++bcol)
                        if (bcol >= a) acc = fma(sT[a * KP + bcol], cw[(d * 8 + g) * CWS + bcol], acc);
                    cf[d][s1] = acc;
                }
            __syncwarp();
#pragma unroll
            for (int d = 0; d < DS; ++d)
#pragma unroll
                for (int s1 = 0; s1 < NS1; ++s1) {
                    cw[(d * 8 + g) * CWS + 4 * s1 + t] = cf[d][s1];
                    ncf[d][s1] = -cf[d][s1];
                }
            __syncwarp();
            if (QUAD) {
                // S = q + 2 s1 + e0 + sum_a c_a (-2 (Vh'(p u~))_a - 2 rv_a + (M c)_a);  x_0 from row 0
                const double e0 = hdr[PFB_HDR_E0(KP)];
                const double* sRv = sM + KP * KP;
#pragma unroll
                for (int d = 0; d < DS; ++d) {
                    const double* cv = cw + (d * 8 + g) * CWS;
                    double part = 0.0, v0c = 0.0;
#pragma unroll
                    for (int s1 = 0; s1 < NS1; ++s1) {
                        const int a = 4 * s1 + t;
                        double mc = 0.0;
#pragma unroll
                        for (int kk = 0; kk < KP; ++kk) mc = fma(sM[a * KP + kk], cv[kk], mc);
                        part = fma(cf[d][s1], mc - 2.0 * (cv[KP + a] + sRv[a]), part);
                        v0c = fma(fr[a], cf[d][s1], v0c);  // Vh[0][a] c_a (row 0: swizzle 0)
                    }
                    part += qsum[d];
                    part += __shfl_xor_sync(0xffffffffu, part, 1);
                    part += __shfl_xor_sync(0xffffffffu, part, 2);
                    v0c += __shfl_xor_sync(0xffffffffu, v0c, 1);
                    v0c += __shfl_xor_sync(0xffffffffu, v0c, 2);
                    const double ut0 = __shfl_sync(0xffffffffu, u0[d], lane & ~3);  // u~_0 lives in lane t = 0
                    macc[d].a = part + e0;
                    macc[d].b = fma(fr[KP], ut0 - v0c, fr[KP + 1]);  // x_0 = a_0 (u~_0 - v_0'c) + mu_0
                }
            }
        }
        if (!QUAD) run_pass(pfb_ic<1>{});

        // ---- per-draw results ---------------------------------------------------------------------
#pragma unroll
        for (int d = 0; d < DS; ++d) {
            double us = unormsq[d];
            us += __shfl_xor_sync(0xffffffffu, us, 1);
            us += __shfl_xor_sync(0xffffffffu, us, 2);
            if (!QUAD) macc[d].group_reduce();
            if (t == 0 && ((actmask >> (2 * d)) & 1u)) {
                double logq = (fma((double)n, PFB_LOG2PI, logdet) + us) / -2.0;
                if (!pd_ok) logq = NAN;
                const int64_t oc = MATERIALIZE ? ocol[d] : (int64_t)slot * K + kd[d];
                if (MATERIALIZE) {
                    if (logp_out) logp_out[oc] = macc[d].finish(n, mp);
                    if (logq_out) logq_out[oc] = logq;
                } else {
                    logp_out[oc] = macc[d].finish(n, mp);
                    logq_out[oc] = logq;
                }
            }
        }
        __syncwarp();
    }
}

static size_t k3_smem_bytes(int KP, int NS, int NW, bool quad) {
    const int RS2 = pfb_rs2_of(KP);
    return (size_t)NS * PFB_K3_RC * RS2 * 8 + (size_t)PF_ZIG_LAYERS * PFB_K3_ZREP * sizeof(uint64_t) +
           (size_t)2 * KP * KP * 8 + (quad ? (size_t)(KP * KP + KP) * 8 : 0) +
           (size_t)NW * PFB_K3_DS * 8 * (quad ? 2 * KP + 2 : KP + 1) * 8 + (size_t)NW * sizeof(pfb_k3_warp_list) +
           (size_t)(NS + 1) * 8;
}

template <int KP, int MODEL>
static cudaError_t launch_k3_m(cudaStream_t st, int n, int K, int nslots, const int32_t* unit_list,
                               const double* FR2, const double* HDR, const uint64_t* seeds,
                               const double* u_host, pfb_model_params mp, double* logp, double* logq,
                               double* draws, int two_pass, pfb_k3_sel sel, pfb_k3_fb fb) {
    if (nslots <= 0) return cudaSuccess;
    if (sel.cnt != nullptr && draws == nullptr) return cudaErrorInvalidValue;
    if (n >= (1 << 20)) return cudaErrorInvalidValue;  // the slow-path queue packs the row into 20 bits
    // every registered family is diagonal-quadratic, so the lean path runs single pass unless
    // the caller asks for the generic two-pass kernel (or wants x written out)
    const bool quad = (draws == nullptr) && !two_pass;
    int dev = 0, nsm = 148, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // warps: enough for K draws, at most 16 (256 draws per sweep)
    int NW = (K + PFB_K3_DS * 8 - 1) / (PFB_K3_DS * 8);
    NW = NW < 1 ? 1 : (NW > PFB_K3_MAXWARPS ? PFB_K3_MAXWARPS : NW);
    if (sel.cnt != nullptr && NW > 4) NW = 4;  // selections are short lists: 64 draws per sweep
    const int DPS = NW * PFB_K3_DS * 8;
    const int S = (K + DPS - 1) / DPS;
    const int C = (pfb_npad8(n) + PFB_K3_RC - 1) / PFB_K3_RC;
    int NS = C;
    while (NS > 1 && k3_smem_bytes(KP, NS, NW, quad) > (size_t)smem_max) --NS;
    const size_t smem = k3_smem_bytes(KP, NS, NW, quad);
    if (smem > (size_t)smem_max) return cudaErrorInvalidConfiguration;
    // split a unit's sweeps over several CTAs only when there are too few units to fill the GPU
    int splits = (4 * nsm + nslots - 1) / nslots;
    splits = splits < 1 ? 1 : (splits > S ? S : splits);
    // column selection: the number of sweeps is per slot and only known on the device; resampling
    // with concentrated weights puts most columns into a few slots, so give every slot enough CTAs for
    // the worst case (sel.cap columns) — the ones without a sweep return at once
    if (sel.cnt != nullptr) {
        splits = (sel.cap + DPS - 1) / DPS;
        splits = splits < 1 ? 1 : (splits > 32 ? 32 : splits);
    }
    const int64_t grid = (int64_t)nslots * splits;
    if (grid > 2147483647LL) return cudaErrorInvalidValue;
    void (*kern)(int, int, int, int, const int32_t*, const double*, const double*, const uint64_t*, const double*,
                 pfb_model_params, double*, double*, double*, pfb_k3_sel, pfb_k3_fb);
    if (sel.cnt != nullptr) {
        kern = pfb_k3_elbo_sample<KP, MODEL, 1, true>;
    } else if constexpr (MODEL == PFB_MODEL_EXTERNAL) {
        kern = pfb_k3_elbo_sample<KP, MODEL, 1>;
    } else {
        kern = draws ? pfb_k3_elbo_sample<KP, MODEL, 1>
                     : (quad ? pfb_k3_elbo_sample<KP, MODEL, 2> : pfb_k3_elbo_sample<KP, MODEL, 0>);
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)grid, NW * 32, smem, st>>>(n, K, splits, NS, unit_list, FR2, HDR, seeds, u_host, mp, logp,
                                                logq, draws, sel, fb);
    return cudaGetLastError();
}

template <int KP>
static cudaError_t launch_k3_k(cudaStream_t st, int model, int n, int K, int nslots,
                               const int32_t* unit_list, const double* FR2, const double* HDR,
                               const uint64_t* seeds, const double* u_host, pfb_model_params mp,
                               double* logp, double* logq, double* draws, int two_pass, pfb_k3_sel sel, pfb_k3_fb fb) {
    switch (model) {
        case PFB_MODEL_ISONORMAL:
            return launch_k3_m<KP, PFB_MODEL_ISONORMAL>(st, n, K, nslots, unit_list, FR2, HDR, seeds, u_host,
                                                        mp, logp, logq, draws, two_pass, sel, fb);
        case PFB_MODEL_FUNNEL:
            return launch_k3_m<KP, PFB_MODEL_FUNNEL>(st, n, K, nslots, unit_list, FR2, HDR, seeds, u_host, mp,
                                                     logp, logq, draws, two_pass, sel, fb);
        case PFB_MODEL_DIAGNORMAL:
            return launch_k3_m<KP, PFB_MODEL_DIAGNORMAL>(st, n, K, nslots, unit_list, FR2, HDR, seeds, u_host,
                                                         mp, logp, logq, draws, two_pass, sel, fb);
        case PFB_MODEL_DENSENORMAL:
        case PFB_MODEL_HLOGISTIC:
        case PFB_MODEL_HOSTCALLBACK:
            if (draws == nullptr) return cudaErrorInvalidValue;  // these families need x written out (K8)
            return launch_k3_m<KP, PFB_MODEL_EXTERNAL>(st, n, K, nslots, unit_list, FR2, HDR, seeds, u_host, mp,
                                                       logp, logq, draws, 1, sel, fb);
    }
    return cudaErrorInvalidValue;
}

// u_host != NULL: parity mode (normals [unit][K][n] supplied by the host).
// draws != NULL: materialise x (mode M / K5), [slot][K][n] i.e. column-major n x K per slot.
extern "C" cudaError_t PFB_K3_ENTRY(cudaStream_t st, int model, int n, int K, int nslots,
                                    const int32_t* unit_list, const double* FR2, const double* HDR,
                                    const uint64_t* seeds, const double* u_host, const double* mp0,
                                    const double* mp1, double mc0, double* logp, double* logq,
                                    double* draws, int two_pass, const int32_t* sel_cnt, const void* sel_list,
                                    int sel_cap, const double* fbX, const double* fbG, const int64_t* fb_off,
                                    const uint64_t* fb_seeds, const int32_t* fb_path_of_slot) {
    pfb_model_params mp{mp0, mp1, mc0};
    pfb_k3_sel sel{sel_cnt, reinterpret_cast<const int2*>(sel_list), sel_cap};
    pfb_k3_fb fb{fbX, fbG, fb_off, fb_seeds, fb_path_of_slot};
    return launch_k3_k<PFB_K3_KP>(st, model, n, K, nslots, unit_list, FR2, HDR, seeds, u_host, mp, logp, logq,
                                  draws, two_pass, sel, fb);
}
