// pfb_common.cuh — shared device helpers and the HBM layout of the engine (sm_100a only).
//
// Layout of one batch in HBM (all FP64 unless noted; "unit" = one (path p, iteration l>=1)):
//   X, G        [n x T]   column-major trajectory points / log-density gradients, T = sum(L_p+1)
//   point_off   [P+1]     first column of each path (int64)
//   seeds       [U]       per-unit UInt64 seed, U = sum(L_p); unit u of path p, iteration l is
//                         u = point_off[p] - p + (l-1), its point column is point_off[p] + l
//   alpha       [n x U]   diagonal of H0 per unit                          (K1 -> K2)
//   hist        [U x J]   point columns of the accepted (s,y) pairs, oldest first (int32)
//   hist_cnt    [U]       J_eff (int32)
//   FR          [U][n][RS]  K2's working rows: RS = KP + 2 doubles per row =
//                         { Vh[i][0..KP), sqrt(alpha_i), mu_i }           (K2 scratch, fit export)
//   FR2         [U][npad8(n)][RS2]  the same record in K3's swizzled tensor-core layout (K2 -> K3)
//   HDR         [U][HS]   unit header: T[KP*KP] row-major, Vc[KP*KP] row-major (upper),
//                         logdet, flag (1 = PD ok), k_eff, e0, (pad), M[KP*KP], rv[KP]
//   logp, logq  [U x K]   per-draw log densities                            (K3 -> K4)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PFB_MAX_KP 24  // history_length <= 12

// padded reflector count: 12 (J <= 6), 20 (J <= 10), 24 (J <= 12); 0 = unsupported
__host__ __device__ inline int pfb_kp_of(int J) { return J <= 6 ? 12 : (J <= 10 ? 20 : (J <= 12 ? 24 : 0)); }
__host__ __device__ inline int pfb_rs_of(int KP) { return KP + 2; }
__host__ __device__ inline int pfb_hs_of(int KP) { return 3 * KP * KP + KP + 8; }
// K3's tensor-core record layout (FR2): per unit pfb_npad8(n) rows of RS2 doubles
//   { Vh[i][0..KP), sqrt(alpha_i), mu_i, p_i, 2 r_i, 0... } (p, r: see PFB_HDR_M), rows >= n all zero, and inside a row the
//   4-double groups XOR-swizzled: logical column c lives at c ^ pfb_swz(i).  pfb_swz takes 4
//   distinct values both over rows {4q..4q+3} and over rows {r, r+2, r+4, r+6}, which makes the
//   two DMMA fragment access patterns of K3 shared-memory bank-conflict free.
__host__ __device__ inline int pfb_rs2_of(int KP) { return KP == 12 ? 16 : 32; }
__host__ __device__ inline int pfb_npad8(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int pfb_swz(int row) {
    return (((row >> 1) & 1) | (((row ^ (row >> 2)) & 1) << 1)) << 2;
}
#define PFB_HDR_LOGDET(KP) (2 * (KP) * (KP))
#define PFB_HDR_FLAG(KP) (2 * (KP) * (KP) + 1)
#define PFB_HDR_KEFF(KP) (2 * (KP) * (KP) + 2)
// statistics of the diagonal-quadratic target family for K3's single-pass mode (written by K2):
//   logp(x) = g(S, x_0),  S = sum_i d_i (x_i - m_i)^2  (iso-normal: d = 1, m = 0; funnel: d_0 = 0,
//   d_i = 1, m = 0; independent normals: d_i = 1 / sd_i^2, m_i = mean_i).  With x = a (u~ - Vh c) + mu,
//   p_i = d_i alpha_i and r_i = d_i a_i (mu_i - m_i) (record slots KP+2 and KP+3, the latter as 2 r_i):
//   S = sum p u~^2 - 2 c'(Vh' (p u~)) + c' M c + 2 sum r u~ - 2 c' rv + e0,
//   M = Vh' diag(p) Vh,  rv = Vh' r,  e0 = sum d (mu - m)^2.
#define PFB_HDR_E0(KP) (2 * (KP) * (KP) + 3)
#define PFB_HDR_M(KP) (2 * (KP) * (KP) + 8)
#define PFB_HDR_RV(KP) (3 * (KP) * (KP) + 8)

#define PFB_LOG2PI 1.8378770664093453

// model family ids (registered device-side target log densities): the
// PFB_MODEL_* macros of include/pfb200.h
#include "../../include/pfb200.h"

// ---- deterministic reductions ---------------------------------------------------------------
__device__ __forceinline__ double pfb_warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Block-wide sum of NV values per thread; result broadcast to every thread.  `scratch` needs
// NV * 32 doubles.  Fixed order: warp butterfly, then butterfly over the warp sums.
template <int NV>
__device__ __forceinline__ void pfb_block_sum(double (&v)[NV], double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int a = 0; a < NV; ++a) v[a] = pfb_warp_sum(v[a]);
    __syncthreads();  // scratch may still be read from a previous call
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < NV; ++a) scratch[a * 32 + warp] = v[a];
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NV; ++a) {
        double t = (lane < nwarp) ? scratch[a * 32 + lane] : 0.0;
        v[a] = pfb_warp_sum(t);
    }
}

// Cheaper block-wide sum of the values selected by `mask` (bit a = v[a] takes part; must be the
// same in every thread): the warps fold their 32 lanes to 8 partials with two shuffle steps, park
// them in shared memory, and then ONE warp per value finishes it (instead of every warp
// butterfly-reducing every value).  ~3x fewer instructions than pfb_block_sum for NV = 12.
// `scratch` needs NV * nwarp * 8 + NV doubles.  Fixed order => deterministic.  Result broadcast.
template <int NV>
__device__ __forceinline__ void pfb_block_sum_fast(double (&v)[NV], double* scratch, unsigned mask) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    double* totals = scratch + NV * nwarp * 8;
#pragma unroll
    for (int a = 0; a < NV; ++a)
        if ((mask >> a) & 1u) {
            v[a] += __shfl_xor_sync(0xffffffffu, v[a], 16);
            v[a] += __shfl_xor_sync(0xffffffffu, v[a], 8);
        }
    __syncthreads();  // scratch may still be read from a previous call
    if (lane < 8) {
#pragma unroll
        for (int a = 0; a < NV; ++a)
            if ((mask >> a) & 1u) scratch[(a * nwarp + warp) * 8 + lane] = v[a];
    }
    __syncthreads();
    for (int a = warp; a < NV; a += nwarp) {
        if ((mask >> a) & 1u) {
            double t = 0.0;
            for (int k = lane; k < nwarp * 8; k += 32) t += scratch[a * nwarp * 8 + k];
            t = pfb_warp_sum(t);
            if (lane == 0) totals[a] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NV; ++a)
        if ((mask >> a) & 1u) v[a] = totals[a];
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) ------------------------------------
__device__ __forceinline__ uint32_t pfb_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void pfb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pfb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pfb_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pfb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pfb_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pfb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(pfb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned),
// completion signalled on `bar` through complete_tx.
__device__ __forceinline__ void pfb_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            pfb_smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(pfb_smem_u32(bar))
        : "memory");
}
