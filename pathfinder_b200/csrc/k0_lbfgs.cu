// K0 lbfgs_trajectories — batched device L-BFGS for the registered closed-form families
// (SURVEY §8 row f1).  Replaces the per-path host optimiser
//   optimize_with_trace (src/optimize.jl:35-59) under _chunk_tmap (src/multipath.jl:190-208)
// by ONE launch: one CTA of PF_LBFGS_T threads per path runs the whole trajectory (pf_lbfgs.h:
// two-loop recursion + strong-Wolfe line search) and writes the points / log-density gradients the
// callback records (src/optimize.jl:94-101) straight into HBM slabs, from where a pack kernel
// gathers them into the contiguous n x T layout K1/K2 read — the trajectories never visit the host.
//
// Every vector (direction, S/Y history, the slab columns) is touched with the fixed ownership
// i = tid + k * PF_LBFGS_T, so after the first pass it is L1/L2 resident; the sequential part is
// the ~20 block-wide reductions per iteration (2 J for the two-loop recursion, 2 per density
// evaluation, 2 for the curvature pair).  The reduction order is the contract of pf_lbfgs.h, which
// the CPU oracle emulates, so trajectories are bit-identical to the oracle's.
#include "pfb_common.cuh"
#include "pf_lbfgs.h"

// NT physical threads per path.  Elementwise work is spread over all of them; the reductions keep
// the contract's shape — PF_LBFGS_T strided partials held by the first PF_LBFGS_T threads, the rest
// contribute exact zeros — so the result does not depend on NT.
template <int NT>
struct pfb_lbfgs_dev_ctx {
    int n;
    double* scratch;  // 64 doubles of shared memory
    template <class F>
    __device__ __forceinline__ void each(F f) {
        for (int i = threadIdx.x; i < n; i += NT) f(i);
    }
    template <class F>
    __device__ __forceinline__ void each_n(int count, F f) {
        for (int i = threadIdx.x; i < count; i += NT) f(i);
    }
    // out[i] = init + sum_j A[i + j nr] v[j]: a thread interleaves up to 4 rows (independent chains,
    // so several loads are in flight) and sweeps the columns once; consecutive rows sit in consecutive
    // threads => every column step is a coalesced read
    __device__ __forceinline__ void matvec_cols(int nr, int nc, const double* __restrict__ A,
                                                const double* __restrict__ v, double init, double* out) {
        for (int i0 = threadIdx.x; i0 < nr; i0 += 4 * NT) {
            double acc[4] = {init, init, init, init};
            const int i1 = i0 + NT, i2 = i0 + 2 * NT, i3 = i0 + 3 * NT;
            const bool h1 = i1 < nr, h2 = i2 < nr, h3 = i3 < nr;
#pragma unroll 4
            for (int j = 0; j < nc; ++j) {
                const double* col = A + (size_t)j * nr;
                const double vj = v[j];
                acc[0] = fma(col[i0], vj, acc[0]);
                if (h1) acc[1] = fma(col[i1], vj, acc[1]);
                if (h2) acc[2] = fma(col[i2], vj, acc[2]);
                if (h3) acc[3] = fma(col[i3], vj, acc[3]);
            }
            out[i0] = acc[0];
            if (h1) out[i1] = acc[1];
            if (h2) out[i2] = acc[2];
            if (h3) out[i3] = acc[3];
        }
    }
    template <class F>
    __device__ __forceinline__ double sum(F f) {
        return sum_n(n, f);
    }
    // With NT > PF_LBFGS_T the reductions read elements that `each` (stride NT) wrote from OTHER
    // threads, so every reduction starts with a barrier; with NT == PF_LBFGS_T a thread reduces
    // exactly the elements it wrote itself.
    __device__ __forceinline__ void pre_reduce() {
        if (NT != PF_LBFGS_T) __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ double sum_n(int count, F f) {
        pre_reduce();
        double acc = 0.0;
        if (threadIdx.x < PF_LBFGS_T)
            for (int i = threadIdx.x; i < count; i += PF_LBFGS_T) acc = f(i, acc);
        double v[1] = {acc};
        pfb_block_sum<1>(v, scratch);
        return v[0];
    }
    template <class F>
    __device__ __forceinline__ void sum2(F f, double& a, double& b) {
        pre_reduce();
        double v[2] = {0.0, 0.0};
        if (threadIdx.x < PF_LBFGS_T)
            for (int i = threadIdx.x; i < n; i += PF_LBFGS_T) f(i, v[0], v[1]);
        pfb_block_sum<2>(v, scratch);
        a = v[0];
        b = v[1];
    }
    template <class F>
    __device__ __forceinline__ double maxv(F f) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < n; i += NT) acc = fmax(acc, f(i));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __syncthreads();
        if (lane == 0) scratch[warp] = acc;
        __syncthreads();
        double t = (lane < NT / 32) ? scratch[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, off));
        return t;
    }
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <int NT>
__global__ void __launch_bounds__(NT)
pfb_k0_lbfgs(pf_lbfgs_model m, pf_lbfgs_opts o, const double* __restrict__ x0, double* X, double* G, double* FX,
             double* ws, int64_t* npoints, int32_t* status, int32_t* nevals) {
    __shared__ double scratch[64];
    const int p = blockIdx.x;
    const size_t slab = (size_t)m.n * (size_t)o.max_points;
    pfb_lbfgs_dev_ctx<NT> c{m.n, scratch};
    int st = 0, nev = 0;
    const size_t zlen = (size_t)(m.nobs > m.n ? m.nobs : m.n);
    double* wsp = ws + (size_t)p * ((size_t)(2 * o.J + 1) * m.n + zlen);
    m.zbuf = wsp + (size_t)(2 * o.J + 1) * m.n;
    const int np = pf_lbfgs_run(c, m, o, x0 + (size_t)p * m.n, X + (size_t)p * slab, G + (size_t)p * slab,
                                FX + (size_t)p * o.max_points, wsp, &st, &nev);
    if (threadIdx.x == 0) {
        npoints[p] = np;
        status[p] = st;
        nevals[p] = nev;
    }
}

// dst[:, t] = slab[:, src[t]] for the points and the gradients; one CTA per packed column
__global__ void pfb_k0_pack(int n, const int64_t* __restrict__ src, const double* __restrict__ sX,
                            const double* __restrict__ sG, double* __restrict__ dX, double* __restrict__ dG) {
    const int64_t t = blockIdx.x;
    const int64_t s = src[t];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        dX[t * n + i] = sX[s * n + i];
        dG[t * n + i] = sG[s * n + i];
    }
}

extern "C" cudaError_t pfb_launch_k0(cudaStream_t st, int family, int n, int nobs, int P, const double* mp0,
                                     const double* mp1, const double* mp2, double mc0, int J, int maxiters, int max_points, double gtol, double ftol,
                                     const double* x0, double* X, double* G, double* FX, double* ws,
                                     int64_t* npoints, int32_t* status, int32_t* nevals) {
    if (P <= 0) return cudaSuccess;
    pf_lbfgs_model m{family, n, mp0, mp1, mc0, nullptr, nobs, mp2};
    pf_lbfgs_opts o{J, maxiters, max_points, gtol, ftol};
    // the GEMM-shaped families stream a matrix per evaluation: more threads = more loads in flight
    if (family == PF_LBFGS_DENSENORMAL || family == PF_LBFGS_HLOGISTIC)
        pfb_k0_lbfgs<512><<<P, 512, 0, st>>>(m, o, x0, X, G, FX, ws, npoints, status, nevals);
    else
        pfb_k0_lbfgs<PF_LBFGS_T><<<P, PF_LBFGS_T, 0, st>>>(m, o, x0, X, G, FX, ws, npoints, status, nevals);
    return cudaGetLastError();
}

extern "C" cudaError_t pfb_launch_k0_pack(cudaStream_t st, int n, int64_t T, const int64_t* src, const double* sX,
                                          const double* sG, double* dX, double* dG) {
    if (T <= 0) return cudaSuccess;
    pfb_k0_pack<<<(unsigned)T, 256, 0, st>>>(n, src, sX, sG, dX, dG);
    return cudaGetLastError();
}
