// pfb_api.cu — the C ABI of libpfb200.so (declared in include/pfb200.h): engine handle,
// device workspace (grow-only, engine-owned), and the orchestration of K1..K7 on one stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/pfb200.h"
#include "pfb_common.cuh"

extern "C" {
cudaError_t pfb_launch_k1(cudaStream_t, int, int, int, double, const double*, const double*, const int64_t*,
                          double*, int32_t*, int32_t*, int64_t*);
cudaError_t pfb_launch_k2(cudaStream_t, int, int, int, int, const double*, const double*, const int32_t*,
                          const double*, const int32_t*, const int32_t*, double*, double*, double*, int,
                          const double*, const double*);
#define PFB_DECL_K3(name)                                                                                  \
    cudaError_t name(cudaStream_t, int, int, int, int, const int32_t*, const double*, const double*,        \
                     const uint64_t*, const double*, const double*, const double*, double, double*, double*, \
                     double*, int, const int32_t*, const void*, int, const double*, const double*,           \
                     const int64_t*, const uint64_t*, const int32_t*);
PFB_DECL_K3(pfb_launch_k3_kp12)
PFB_DECL_K3(pfb_launch_k3_kp20)
PFB_DECL_K3(pfb_launch_k3_kp24)
cudaError_t pfb_launch_k4(cudaStream_t, int, int, int64_t, const int64_t*, const double*, const double*, double*,
                          double*, int64_t*, int32_t*, int32_t*);
cudaError_t pfb_launch_k1_range(cudaStream_t, int, int, int, int, double, const double*, const double*, const int64_t*,
                                double*, int32_t*, int32_t*, int64_t*);
cudaError_t pfb_launch_k2_range(cudaStream_t, int, int, int, int, int, const double*, const double*, const int32_t*,
                                const double*, const int32_t*, const int32_t*, double*, double*, double*, int,
                                const double*, const double*);
int pfb_k2_uses_smem_panel(int KP, int n);
// generic (runtime-width) kernels for history_length > 12 (kg_generic_history.cu)
size_t pfb_kg_k2_workspace_doubles(int KP, int grid);
int pfb_kg_k2_grid(int U);
cudaError_t pfb_launch_k2g(cudaStream_t, int, int, int, int, int, const double*, const double*, const int32_t*,
                           const double*, const int32_t*, const int32_t*, double*, double*, double*);
size_t pfb_kg_k3_scratch_doubles(int KP, int n, int nslots);
cudaError_t pfb_launch_k3g(cudaStream_t, int, int, int, int, int, const int32_t*, const double*, const double*,
                           const uint64_t*, const double*, const double*, const double*, double, double*, double*,
                           double*, const int32_t*, const void*, int, const double*, const double*, const int64_t*,
                           const uint64_t*, const int32_t*, double*);
cudaError_t pfb_launch_kg_gather_fit(cudaStream_t, int, int, int, const int32_t*, const double*, const double*,
                                     const double*, const int32_t*, double*, double*, double*, double*, double*,
                                     double*, int32_t*);
cudaError_t pfb_launch_k8g_dense(cudaStream_t, int, int64_t, int, const int32_t*, const double*, const double*,
                                 const double*, double, double*);
cudaError_t pfb_launch_k8g_logistic(cudaStream_t, int, int, int64_t, int, const int32_t*, const double*,
                                    const double*, const double*, double*);
cudaError_t pfb_launch_k0(cudaStream_t, int, int, int, int, const double*, const double*, const double*, double, int, int,
                          int, double,
                          double, const double*, double*, double*, double*, double*, int64_t*, int32_t*, int32_t*);
cudaError_t pfb_launch_k0_pack(cudaStream_t, int, int64_t, const int64_t*, const double*, const double*, double*,
                               double*);
size_t pfb_psis_scalars_size();
cudaError_t pfb_launch_k7r_bin(cudaStream_t, int, int, int, int64_t, const int64_t*, int32_t*, void*);
cudaError_t pfb_k7b_temp_bytes(int, size_t*);
cudaError_t pfb_launch_k7b(cudaStream_t, int, int, int, uint64_t, int, const double*, const double*, uint64_t*,
                           uint64_t*, int32_t*, int32_t*, void*, size_t, int64_t*, int64_t*, double*);
size_t pfb_k6_workspace_bytes(int);
cudaError_t pfb_launch_k6(cudaStream_t, int, int, int, const double*, const double*, const double*, double*,
                          double*, uint64_t*, void*, void*, size_t);
cudaError_t pfb_launch_k7(cudaStream_t, int, int, int, uint64_t, int, const uint64_t*, const void*,
                          const double*, int64_t*, int64_t*, double*);
}

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct psis_scalars_host {
    double pareto_k, lse, sigma, logu;
    uint64_t Z;
    int64_t tail_len, smoothed;
    uint64_t maxkey;
};

}  // namespace

struct pfb_engine {
    pfb_config cfg;
    int KP = 12;
    bool generic = false;  // history_length > 12: the runtime-width kernels K2g / K3g (KP = 2 J)
    DevBuf dKgWs, dKgScratch;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[7] = {};
    std::string err;
    // model
    int model = -1, model_n = 0;
    DevBuf dModel;
    double model_c0 = 0.0;
    int model_nobs = 0;                 // HLOGISTIC
    DevBuf dGenX, dIota, dTopSeeds;
    // batch state
    int n = 0, P = 0, K = 0;
    int64_t T = 0, U = 0;
    bool have_batch = false, ran = false, have_normals = false;
    int poolK = 0;  // draws per path in the device pool (K, or the count of the last pfb_draw_from_fits)
    int poolP = 0;  // paths in the device pool (P, unless pfb_pool_set installed a pool assembled on the host)
    bool pool_ready = false;  // the pool's DRAWS are materialised (its logp / logq always are after a run)
    DevBuf dSelCnt, dSelList;
    // failed paths (src/singlepath.jl:224-228): fresh draws from fit_distributions[fit_iteration + 1]
    std::vector<uint64_t> fb_seeds_user;  // pfb_set_fallback_seeds (consumed by the next batch)
    std::vector<uint64_t> fb_seeds;       // one per path of the current batch
    DevBuf dFbSeeds, dPoolSeeds, dFbUnits, dFbPaths, dFbLogp, dFbLogq, dFbPairs, dTopFb;
    const uint64_t* pool_seeds = nullptr;  // unit-indexed seeds of the pool's draws (dSeeds, or dPoolSeeds)
    int n_failed = 0;
    // lazy failure resolution: pfb_batch_run leaves the success flags / best units on their way to this
    // page-locked pair and returns without synchronising; the first consumer of the pool resolves them
    bool fail_pending = false;
    cudaEvent_t ev_fail = nullptr;
    int32_t* hSuccBu = nullptr;  // [2 x cap]: success flags, best units
    int hSuccBuCap = 0;
    // multi-GPU (pfb_comm_init): NCCL communicator, the all-gathered log densities
    void* comm = nullptr;
    int comm_world = 1, comm_rank = 0;
    DevBuf dGLogp, dGLogq;
    int launches = 0;
    std::vector<int64_t> h_off;
    DevBuf dX, dG, dOff, dSeeds, dUnitCol, dNormals;
    DevBuf dAlpha, dHist, dHistCnt, dRej, dFR, dFR2, dHDR, dLogp, dLogq, dElbo, dSe, dBestIter, dBestUnit, dSucc;
    DevBuf dPool, dPoolLogp, dPoolLogq, dAllDraws;
    DevBuf dFitMu, dFitAlpha, dFitVh, dFitT, dFitVc, dFitLogdet, dFitJeff;
    // host-callback target (row f2): pinned staging, copy stream
    pfb_logp_callback host_cb = nullptr;
    void* host_user = nullptr;
    double* hX[2] = {nullptr, nullptr};
    size_t hX_cap = 0;
    double* hLp = nullptr;
    size_t hLp_cap = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t hc_k[2] = {}, hc_c[2] = {};
    double host_cb_s = 0.0;  // seconds spent inside the callback during the last run
    // pipelined upload of pfb_elbo_batch: path groups, one event per group
    static constexpr int kUpGroups = 4;
    cudaEvent_t up_ev[kUpGroups] = {};
    cudaEvent_t k1_ev[kUpGroups] = {};
    cudaStream_t k1_stream = nullptr;  // K1 is a per-path latency chain: its groups run beside K2 / the copies
    int up_ngroups = 0;               // 0: the trajectories are already resident
    int up_p[kUpGroups + 1] = {};     // path boundaries of the groups
    // device L-BFGS (K0): trajectory slabs [n x max_points] per path
    DevBuf dLbX0, dLbX, dLbG, dLbFX, dLbWs, dLbNp, dLbSt, dLbNev, dLbSrc;
    std::vector<int64_t> lb_np;
    int lb_P = 0, lb_n = 0, lb_maxpts = 0;
    bool lb_ok = false;
    float lb_ms = 0.f;
    cudaEvent_t lb_ev[2] = {};
    // psis
    DevBuf dLogw, dW, dCum, dScal, dInds, dIds, dOutDraws, dTmpLogr, dTmpPool, dSortWork, dSortTmp;
};

#define PFB_D2H(dst, src, bytes)                                                                   \
    do {                                                                                           \
        if ((dst) && (bytes) > 0)                                                                  \
            PFB_CUDA(h, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, st));        \
    } while (0)

#define PFB_FAIL(h, code, msg)     \
    do {                           \
        (h)->err = (msg);          \
        return (code);             \
    } while (0)

#define PFB_CUDA(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                        \
            return (int)e_ > 0 ? (int)e_ : 1;                                                     \
        }                                                                                         \
    } while (0)

static thread_local std::string g_create_err;

extern "C" int pfb_create(pfb_handle* out, const pfb_config* cfg) {
    if (!out || !cfg) return PFB_ERR_ARG;
    *out = nullptr;
    if (cfg->history_length < 1 || cfg->ndraws_elbo < 1) {
        g_create_err = "history_length and ndraws_elbo must be positive";
        return PFB_ERR_ARG;
    }
    int kp = pfb_kp_of(cfg->history_length);
    const bool generic = (kp == 0);
    if (generic) {
        // any history length like the reference (src/inverse_hessian.jl:25): beyond 12 the generic kernels
        // take over (K1's ring buffer holds 64 pairs)
        if (cfg->history_length > 64) {
            g_create_err = "history_length > 64 is not supported";
            return PFB_ERR_UNSUPPORTED;
        }
        kp = 2 * cfg->history_length;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return e != cudaSuccess ? (int)e : 100;  // cudaErrorNoDevice
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        g_create_err = "bad device ordinal";
        return PFB_ERR_ARG;
    }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        return (int)e;
    }
    pfb_engine* h = new pfb_engine();
    h->cfg = *cfg;
    if (h->cfg.eps == 0.0) h->cfg.eps = 1e-12;
    h->KP = kp;
    h->generic = generic;
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        delete h;
        return (int)e;
    }
    for (auto& ev : h->ev) cudaEventCreate(&ev);
    for (auto& ev : h->lb_ev) cudaEventCreate(&ev);
    *out = h;
    return PFB_OK;
}

extern "C" int pfb_comm_destroy(pfb_handle h);
extern "C" int pfb_destroy(pfb_handle h) {
    if (!h) return PFB_OK;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    if (h->comm) pfb_comm_destroy(h);
    DevBuf* bufs[] = {&h->dModel, &h->dX, &h->dG, &h->dOff, &h->dSeeds, &h->dUnitCol, &h->dNormals, &h->dAlpha,
                      &h->dHist, &h->dHistCnt, &h->dRej, &h->dFR, &h->dFR2, &h->dHDR, &h->dLogp, &h->dLogq, &h->dElbo,
                      &h->dSe, &h->dBestIter, &h->dBestUnit, &h->dSucc, &h->dPool, &h->dPoolLogp,
                      &h->dPoolLogq, &h->dAllDraws, &h->dFitMu, &h->dFitAlpha, &h->dFitVh, &h->dFitT,
                      &h->dFitVc, &h->dFitLogdet, &h->dFitJeff, &h->dLogw, &h->dW, &h->dCum, &h->dScal,
                      &h->dInds, &h->dIds, &h->dOutDraws, &h->dTmpLogr, &h->dTmpPool, &h->dGenX, &h->dIota, &h->dTopSeeds,
                      &h->dLbX0, &h->dLbX, &h->dLbG, &h->dLbFX, &h->dLbWs, &h->dLbNp, &h->dLbSt, &h->dLbNev, &h->dLbSrc, &h->dSortWork, &h->dSortTmp, &h->dSelCnt, &h->dSelList,
                      &h->dFbSeeds, &h->dPoolSeeds, &h->dFbUnits, &h->dFbPaths, &h->dFbLogp, &h->dFbLogq, &h->dFbPairs, &h->dTopFb, &h->dGLogp, &h->dGLogq,
                      &h->dKgWs, &h->dKgScratch};
    for (auto* b : bufs) b->release();
    for (int i = 0; i < 2; ++i) {
        if (h->hX[i]) cudaFreeHost(h->hX[i]);
        if (h->hc_k[i]) cudaEventDestroy(h->hc_k[i]);
        if (h->hc_c[i]) cudaEventDestroy(h->hc_c[i]);
    }
    if (h->hLp) cudaFreeHost(h->hLp);
    if (h->hSuccBu) cudaFreeHost(h->hSuccBu);
    if (h->ev_fail) cudaEventDestroy(h->ev_fail);
    for (auto& ev : h->up_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->k1_ev)
        if (ev) cudaEventDestroy(ev);
    if (h->k1_stream) cudaStreamDestroy(h->k1_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto& ev : h->ev) cudaEventDestroy(ev);
    for (auto& ev : h->lb_ev) cudaEventDestroy(ev);
    cudaStreamDestroy(h->stream);
    delete h;
    return PFB_OK;
}

extern "C" const char* pfb_last_error(pfb_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }
extern "C" int pfb_kp(pfb_handle h) { return h ? h->KP : 0; }

extern "C" int pfb_register_model(pfb_handle h, int family, int n, const double* blob, size_t ndoubles) {
    if (!h) return PFB_ERR_ARG;
    if (n < 1) PFB_FAIL(h, PFB_ERR_ARG, "model dimension must be positive");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    switch (family) {
        case PFB_MODEL_ISONORMAL:
        case PFB_MODEL_FUNNEL:
            break;
        case PFB_MODEL_DIAGNORMAL: {
            if (!blob || ndoubles != (size_t)2 * n) PFB_FAIL(h, PFB_ERR_SHAPE, "DIAGNORMAL blob = {mean[n], sd[n]}");
            std::vector<double> tmp(2 * (size_t)n);
            double c0 = -0.5 * n * PFB_LOG2PI;
            for (int i = 0; i < n; ++i) {
                if (!(blob[n + i] > 0.0)) PFB_FAIL(h, PFB_ERR_ARG, "DIAGNORMAL sd must be positive");
                tmp[i] = blob[i];
                tmp[n + i] = 1.0 / blob[n + i];
                c0 -= log(blob[n + i]);
            }
            PFB_CUDA(h, h->dModel.ensure(tmp.size() * 8));
            PFB_CUDA(h, cudaMemcpyAsync(h->dModel.p, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice, h->stream));
            PFB_CUDA(h, cudaStreamSynchronize(h->stream));
            h->model_c0 = c0;
            break;
        }
        case PFB_MODEL_DENSENORMAL: {
            const size_t nn = (size_t)n;
            if (!blob || ndoubles != nn + nn * nn) PFB_FAIL(h, PFB_ERR_SHAPE, "DENSENORMAL blob = {m[n], P[n x n]}");
            // device: { m[n], P m [n], P[n x n] };  c0 = m' P m
            std::vector<double> pm(nn, 0.0);
            const double* m = blob;
            const double* P = blob + nn;
            for (size_t j = 0; j < nn; ++j) {
                const double mj = m[j];
                const double* col = P + j * nn;
                for (size_t i = 0; i < nn; ++i) pm[i] += col[i] * mj;
            }
            double mPm = 0.0;
            for (size_t i = 0; i < nn; ++i) mPm += m[i] * pm[i];
            PFB_CUDA(h, h->dModel.ensure((2 * nn + nn * nn) * 8));
            double* d = h->dModel.as<double>();
            PFB_CUDA(h, cudaMemcpyAsync(d, m, nn * 8, cudaMemcpyHostToDevice, h->stream));
            PFB_CUDA(h, cudaMemcpyAsync(d + nn, pm.data(), nn * 8, cudaMemcpyHostToDevice, h->stream));
            PFB_CUDA(h, cudaMemcpyAsync(d + 2 * nn, P, nn * nn * 8, cudaMemcpyHostToDevice, h->stream));
            PFB_CUDA(h, cudaStreamSynchronize(h->stream));
            h->model_c0 = mPm;
            break;
        }
        case PFB_MODEL_HLOGISTIC: {
            if (n < 3 || !blob || ndoubles < 1) PFB_FAIL(h, PFB_ERR_SHAPE, "HLOGISTIC needs n >= 3 and a blob");
            const long long nobs = (long long)blob[0];
            const size_t nb = (size_t)n - 2;
            if (nobs < 1 || ndoubles != 1 + (size_t)nobs * nb + (size_t)nobs)
                PFB_FAIL(h, PFB_ERR_SHAPE, "HLOGISTIC blob = {nobs, X[nobs x (n-2)], y[nobs]}");
            // device: { X[nobs x (n-2)], y[nobs], X'[(n-2) x nobs] } (the transposed copy gives K0's
            // X'r sweep coalesced reads)
            const size_t nx = (size_t)nobs * nb;
            std::vector<double> xt(nx);
            for (size_t j = 0; j < nb; ++j)
                for (size_t k = 0; k < (size_t)nobs; ++k) xt[j + k * nb] = blob[1 + k + j * (size_t)nobs];
            PFB_CUDA(h, h->dModel.ensure((ndoubles - 1 + nx) * 8));
            PFB_CUDA(h, cudaMemcpyAsync(h->dModel.p, blob + 1, (ndoubles - 1) * 8, cudaMemcpyHostToDevice, h->stream));
            PFB_CUDA(h, cudaMemcpyAsync(h->dModel.as<double>() + (ndoubles - 1), xt.data(), nx * 8,
                                        cudaMemcpyHostToDevice, h->stream));
            PFB_CUDA(h, cudaStreamSynchronize(h->stream));
            h->model_nobs = (int)nobs;
            break;
        }
        default:
            PFB_FAIL(h, PFB_ERR_UNSUPPORTED, "unknown model family");
    }
    h->model = family;
    h->model_n = n;
    return PFB_OK;
}

// Validates the batch shape, sizes every device buffer (grow-only) and uploads the small index
// arrays (offsets, unit -> point column).  Shared by pfb_batch_upload and pfb_batch_from_lbfgs.
static int batch_prepare(pfb_engine* h, int n, int P, const int64_t* offsets) {
    if (n < 1 || P < 0 || !offsets) PFB_FAIL(h, PFB_ERR_ARG, "bad n / P / offsets");
    if (h->model < 0) PFB_FAIL(h, PFB_ERR_STATE, "no model registered");
    if (h->model_n != n) PFB_FAIL(h, PFB_ERR_SHAPE, "dimension differs from the registered model's");
    if (offsets[0] != 0) PFB_FAIL(h, PFB_ERR_ARG, "offsets[0] must be 0");
    for (int p = 0; p < P; ++p)
        if (offsets[p + 1] < offsets[p] + 1) PFB_FAIL(h, PFB_ERR_SHAPE, "every path needs at least one point");
    const int64_t T = offsets[P], U = T - P;
    if (T > 2147483647LL) PFB_FAIL(h, PFB_ERR_UNSUPPORTED, "too many trajectory points");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    const int K = h->cfg.ndraws_elbo, J = h->cfg.history_length, KP = h->KP;
    h->n = n; h->P = P; h->K = K; h->T = T; h->U = U;
    h->h_off.assign(offsets, offsets + P + 1);
    h->have_batch = false; h->ran = false; h->fail_pending = false;
    std::vector<int32_t> unit_col((size_t)U);
    for (int p = 0; p < P; ++p) {
        const int64_t L = offsets[p + 1] - offsets[p] - 1;
        for (int64_t l = 1; l <= L; ++l) unit_col[(size_t)(offsets[p] - p + l - 1)] = (int32_t)(offsets[p] + l);
    }
    const size_t nT = (size_t)n * (size_t)T, nU = (size_t)n * (size_t)U;
    PFB_CUDA(h, h->dX.ensure(nT * 8 + 8));
    PFB_CUDA(h, h->dG.ensure(nT * 8 + 8));
    PFB_CUDA(h, h->dOff.ensure((size_t)(P + 1) * 8));
    PFB_CUDA(h, h->dSeeds.ensure((size_t)U * 8 + 8));
    PFB_CUDA(h, h->dUnitCol.ensure((size_t)U * 4 + 8));
    PFB_CUDA(h, h->dAlpha.ensure(nU * 8 + 8));
    PFB_CUDA(h, h->dHist.ensure((size_t)U * J * 4 + 8));
    PFB_CUDA(h, h->dHistCnt.ensure((size_t)U * 4 + 8));
    PFB_CUDA(h, h->dRej.ensure((size_t)P * 8 + 8));
    if (h->generic) {
        PFB_CUDA(h, h->dFR.ensure(nU * (size_t)pfb_rs_of(KP) * 8 + 16));  // the generic record layout
        PFB_CUDA(h, h->dKgWs.ensure(pfb_kg_k2_workspace_doubles(KP, pfb_kg_k2_grid((int)U)) * 8 + 16));
    } else {
        if (!pfb_k2_uses_smem_panel(KP, n))  // K2's global-memory workspace (large n only)
            PFB_CUDA(h, h->dFR.ensure(nU * (size_t)pfb_rs_of(KP) * 8 + 16));
        PFB_CUDA(h, h->dFR2.ensure((size_t)pfb_npad8(n) * (size_t)U * (size_t)pfb_rs2_of(KP) * 8 + 16));
    }
    PFB_CUDA(h, h->dHDR.ensure((size_t)U * pfb_hs_of(KP) * 8 + 8));
    PFB_CUDA(h, h->dLogp.ensure((size_t)U * K * 8 + 8));
    PFB_CUDA(h, h->dLogq.ensure((size_t)U * K * 8 + 8));
    PFB_CUDA(h, h->dElbo.ensure((size_t)U * 8 + 8));
    PFB_CUDA(h, h->dSe.ensure((size_t)U * 8 + 8));
    PFB_CUDA(h, h->dBestIter.ensure((size_t)P * 8 + 8));
    PFB_CUDA(h, h->dBestUnit.ensure((size_t)P * 4 + 8));
    PFB_CUDA(h, h->dSucc.ensure((size_t)P * 4 + 8));
    PFB_CUDA(h, h->dPool.ensure((size_t)n * K * (size_t)P * 8 + 8));
    PFB_CUDA(h, h->dPoolLogp.ensure((size_t)K * P * 8 + 8));
    PFB_CUDA(h, h->dPoolLogq.ensure((size_t)K * P * 8 + 8));
    if (h->cfg.materialize_all) PFB_CUDA(h, h->dAllDraws.ensure(nU * (size_t)K * 8 + 8));
    cudaStream_t st = h->stream;
    PFB_CUDA(h, cudaMemcpyAsync(h->dOff.p, offsets, (size_t)(P + 1) * 8, cudaMemcpyHostToDevice, st));
    if (U > 0)
        PFB_CUDA(h, cudaMemcpyAsync(h->dUnitCol.p, unit_col.data(), (size_t)U * 4, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));  // unit_col is a stack-owned staging vector
    return PFB_OK;
}

extern "C" int pfb_batch_upload(pfb_handle h, int n, int P, const int64_t* offsets, const double* positions,
                                const double* gradients, const uint64_t* seeds, const double* normals) {
    if (!h) return PFB_ERR_ARG;
    if (n >= 1 && P >= 0 && offsets) {
        const int64_t T = offsets[P], U = T - P;
        if (T > 0 && (!positions || !gradients)) PFB_FAIL(h, PFB_ERR_ARG, "positions / gradients are NULL");
        if (U > 0 && !seeds) PFB_FAIL(h, PFB_ERR_ARG, "seeds is NULL");
    }
    int rc = batch_prepare(h, n, P, offsets);
    if (rc) return rc;
    const int64_t T = h->T, U = h->U;
    const int K = h->K;
    const size_t nT = (size_t)n * (size_t)T, nU = (size_t)n * (size_t)U;
    cudaStream_t st = h->stream;
    if (T > 0) {
        PFB_CUDA(h, cudaMemcpyAsync(h->dX.p, positions, nT * 8, cudaMemcpyHostToDevice, st));
        PFB_CUDA(h, cudaMemcpyAsync(h->dG.p, gradients, nT * 8, cudaMemcpyHostToDevice, st));
    }
    if (U > 0) PFB_CUDA(h, cudaMemcpyAsync(h->dSeeds.p, seeds, (size_t)U * 8, cudaMemcpyHostToDevice, st));
    h->have_normals = (normals != nullptr);
    if (normals && U > 0) {
        PFB_CUDA(h, h->dNormals.ensure(nU * (size_t)K * 8));
        PFB_CUDA(h, cudaMemcpyAsync(h->dNormals.p, normals, nU * (size_t)K * 8, cudaMemcpyHostToDevice, st));
    }
    PFB_CUDA(h, cudaStreamSynchronize(st));
    h->have_batch = true;
    return PFB_OK;
}

// ---- K0: batched device L-BFGS ---------------------------------------------
static bool model_has_device_lbfgs(const pfb_engine* h) {
    return h->model == PFB_MODEL_ISONORMAL || h->model == PFB_MODEL_FUNNEL || h->model == PFB_MODEL_DIAGNORMAL ||
           h->model == PFB_MODEL_DENSENORMAL || h->model == PFB_MODEL_HLOGISTIC;
}

extern "C" int pfb_lbfgs_batch(pfb_handle h, int n, int P, const double* x0, const pfb_lbfgs_opts* o,
                               int64_t* npoints, int32_t* status, int32_t* nevals) {
    if (!h) return PFB_ERR_ARG;
    if (n < 1 || P < 0 || (P > 0 && !x0) || !o) PFB_FAIL(h, PFB_ERR_ARG, "bad n / P / x0 / opts");
    if (h->model < 0) PFB_FAIL(h, PFB_ERR_STATE, "no model registered");
    if (h->model_n != n) PFB_FAIL(h, PFB_ERR_SHAPE, "dimension differs from the registered model's");
    if (!model_has_device_lbfgs(h))
        PFB_FAIL(h, PFB_ERR_UNSUPPORTED, "device L-BFGS covers the registered device-side families; a host-callback "
                                         "target is optimised on the host");
    if (o->maxiters < 0 || o->max_points < 1) PFB_FAIL(h, PFB_ERR_ARG, "maxiters >= 0 and max_points >= 1 required");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const int J = h->cfg.history_length;
    const int maxpts = std::min(o->max_points, o->maxiters + 1);
    const size_t slab = (size_t)n * (size_t)maxpts * (size_t)P;
    h->lb_ok = false;
    PFB_CUDA(h, h->dLbX0.ensure((size_t)n * P * 8 + 8));
    PFB_CUDA(h, h->dLbX.ensure(slab * 8 + 8));
    PFB_CUDA(h, h->dLbG.ensure(slab * 8 + 8));
    PFB_CUDA(h, h->dLbFX.ensure((size_t)maxpts * P * 8 + 8));
    const size_t zlen = (size_t)std::max(n, h->model == PFB_MODEL_HLOGISTIC ? h->model_nobs : 0);
    PFB_CUDA(h, h->dLbWs.ensure(((size_t)(2 * J + 1) * n + zlen) * P * 8 + 8));
    PFB_CUDA(h, h->dLbNp.ensure((size_t)P * 8 + 8));
    PFB_CUDA(h, h->dLbSt.ensure((size_t)P * 4 + 8));
    PFB_CUDA(h, h->dLbNev.ensure((size_t)P * 4 + 8));
    h->lb_np.assign((size_t)P, 0);
    h->lb_P = P; h->lb_n = n; h->lb_maxpts = maxpts;
    if (P == 0) {
        h->lb_ok = true;
        return PFB_OK;
    }
    PFB_CUDA(h, cudaMemcpyAsync(h->dLbX0.p, x0, (size_t)n * P * 8, cudaMemcpyHostToDevice, st));
    // DIAGNORMAL blob on the device: { mean[n], 1/sd[n] };  DENSENORMAL: { m[n], P m[n], P[n x n] }
    //   HLOGISTIC: { X[nobs x (n-2)], y[nobs] }
    const double* mp0 = (h->model == PFB_MODEL_DIAGNORMAL || h->model == PFB_MODEL_DENSENORMAL ||
                         h->model == PFB_MODEL_HLOGISTIC) ? h->dModel.as<double>() : nullptr;
    const double* mp1 = !mp0 ? nullptr
                        : (h->model == PFB_MODEL_DENSENORMAL ? mp0 + 2 * (size_t)n
                           : (h->model == PFB_MODEL_HLOGISTIC ? mp0 + (size_t)h->model_nobs * (n - 2) : mp0 + n));
    const int nobs = h->model == PFB_MODEL_HLOGISTIC ? h->model_nobs : 0;
    const double* mp2 = h->model == PFB_MODEL_HLOGISTIC ? mp1 + nobs : nullptr;  // X' follows y
    const double mc0 = h->model == PFB_MODEL_HLOGISTIC ? (-0.5 * n * PFB_LOG2PI - log(2.5)) : h->model_c0;
    PFB_CUDA(h, cudaEventRecord(h->lb_ev[0], st));
    PFB_CUDA(h, pfb_launch_k0(st, h->model, n, nobs, P, mp0, mp1, mp2, mc0, J, o->maxiters, maxpts,
                              o->gtol, o->ftol, h->dLbX0.as<double>(), h->dLbX.as<double>(), h->dLbG.as<double>(),
                              h->dLbFX.as<double>(), h->dLbWs.as<double>(), h->dLbNp.as<int64_t>(),
                              h->dLbSt.as<int32_t>(), h->dLbNev.as<int32_t>()));
    PFB_CUDA(h, cudaEventRecord(h->lb_ev[1], st));
    PFB_CUDA(h, cudaMemcpyAsync(h->lb_np.data(), h->dLbNp.p, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    if (status) PFB_CUDA(h, cudaMemcpyAsync(status, h->dLbSt.p, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    if (nevals) PFB_CUDA(h, cudaMemcpyAsync(nevals, h->dLbNev.p, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));
    PFB_CUDA(h, cudaEventElapsedTime(&h->lb_ms, h->lb_ev[0], h->lb_ev[1]));
    if (npoints) memcpy(npoints, h->lb_np.data(), (size_t)P * 8);
    h->lb_ok = true;
    return PFB_OK;
}

extern "C" int pfb_batch_from_lbfgs(pfb_handle h, const uint64_t* seeds) {
    if (!h) return PFB_ERR_ARG;
    if (!h->lb_ok) PFB_FAIL(h, PFB_ERR_STATE, "pfb_lbfgs_batch has not been called");
    const int n = h->lb_n, P = h->lb_P;
    std::vector<int64_t> off((size_t)P + 1, 0);
    for (int p = 0; p < P; ++p) off[(size_t)p + 1] = off[(size_t)p] + h->lb_np[(size_t)p];
    const int64_t T = off[(size_t)P], U = T - P;
    if (U > 0 && !seeds) PFB_FAIL(h, PFB_ERR_ARG, "seeds is NULL");
    int rc = batch_prepare(h, n, P, off.data());
    if (rc) return rc;
    cudaStream_t st = h->stream;
    std::vector<int64_t> src((size_t)T);
    for (int p = 0; p < P; ++p)
        for (int64_t l = 0; l < h->lb_np[(size_t)p]; ++l)
            src[(size_t)(off[(size_t)p] + l)] = (int64_t)p * h->lb_maxpts + l;
    PFB_CUDA(h, h->dLbSrc.ensure((size_t)T * 8 + 8));
    if (T > 0) {
        PFB_CUDA(h, cudaMemcpyAsync(h->dLbSrc.p, src.data(), (size_t)T * 8, cudaMemcpyHostToDevice, st));
        PFB_CUDA(h, pfb_launch_k0_pack(st, n, T, h->dLbSrc.as<int64_t>(), h->dLbX.as<double>(), h->dLbG.as<double>(),
                                       h->dX.as<double>(), h->dG.as<double>()));
    }
    if (U > 0) PFB_CUDA(h, cudaMemcpyAsync(h->dSeeds.p, seeds, (size_t)U * 8, cudaMemcpyHostToDevice, st));
    h->have_normals = false;
    PFB_CUDA(h, cudaStreamSynchronize(st));  // src is a stack-owned staging vector
    h->have_batch = true;
    return PFB_OK;
}

extern "C" int pfb_lbfgs_download(pfb_handle h, double* positions, double* gradients, double* log_densities) {
    if (!h) return PFB_ERR_ARG;
    if (!h->lb_ok) PFB_FAIL(h, PFB_ERR_STATE, "pfb_lbfgs_batch has not been called");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const size_t n = (size_t)h->lb_n, mp = (size_t)h->lb_maxpts;
    int64_t off = 0;
    for (int p = 0; p < h->lb_P; ++p) {
        const size_t np = (size_t)h->lb_np[(size_t)p];
        if (positions)
            PFB_CUDA(h, cudaMemcpyAsync(positions + (size_t)off * n, h->dLbX.as<double>() + (size_t)p * mp * n,
                                        np * n * 8, cudaMemcpyDeviceToHost, st));
        if (gradients)
            PFB_CUDA(h, cudaMemcpyAsync(gradients + (size_t)off * n, h->dLbG.as<double>() + (size_t)p * mp * n,
                                        np * n * 8, cudaMemcpyDeviceToHost, st));
        if (log_densities)
            PFB_CUDA(h, cudaMemcpyAsync(log_densities + off, h->dLbFX.as<double>() + (size_t)p * mp, np * 8,
                                        cudaMemcpyDeviceToHost, st));
        off += (int64_t)np;
    }
    PFB_CUDA(h, cudaStreamSynchronize(st));
    return PFB_OK;
}

extern "C" int pfb_lbfgs_ms(pfb_handle h, double* ms) {
    if (!h || !ms) return PFB_ERR_ARG;
    if (!h->lb_ok) PFB_FAIL(h, PFB_ERR_STATE, "pfb_lbfgs_batch has not been called");
    *ms = h->lb_ms;
    return PFB_OK;
}

// K2 over units [u0, u0 + cnt): the tensor-core kernel, or the generic one for wide histories
static cudaError_t launch_k2_any(pfb_engine* h, int u0, int cnt, int k2_model, const double* mp0, const double* mp1) {
    if (h->generic)
        return pfb_launch_k2g(h->stream, h->KP, h->n, u0, cnt, h->cfg.history_length, h->dX.as<double>(),
                              h->dG.as<double>(), h->dUnitCol.as<int32_t>(), h->dAlpha.as<double>(),
                              h->dHist.as<int32_t>(), h->dHistCnt.as<int32_t>(), h->dFR.as<double>(),
                              h->dHDR.as<double>(), h->dKgWs.as<double>());
    return pfb_launch_k2_range(h->stream, h->KP, h->n, u0, cnt, h->cfg.history_length, h->dX.as<double>(),
                               h->dG.as<double>(), h->dUnitCol.as<int32_t>(), h->dAlpha.as<double>(),
                               h->dHist.as<int32_t>(), h->dHistCnt.as<int32_t>(), h->dFR.as<double>(),
                               h->dHDR.as<double>(), h->dFR2.as<double>(), k2_model, mp0, mp1);
}

static bool model_is_external(const pfb_engine* h) {
    return h->model == PFB_MODEL_DENSENORMAL || h->model == PFB_MODEL_HLOGISTIC ||
           h->model == PFB_MODEL_HOSTCALLBACK;
}

// fb_seeds != NULL: slots whose unit is < 0 draw from the identity fit of iteration 0 (one seed per path;
// fb_paths maps slots to paths, NULL = identity) instead of writing NaN
static cudaError_t launch_k3(pfb_engine* h, int nslots, const int32_t* unit_list, double* logp, double* logq,
                             double* draws, int K_over = 0, const uint64_t* seeds_over = nullptr,
                             const int32_t* sel_cnt = nullptr, const void* sel_list = nullptr, int sel_cap = 0,
                             const uint64_t* fb_seeds = nullptr, const int32_t* fb_paths = nullptr);

extern "C" int pfb_register_host_model(pfb_handle h, int n, pfb_logp_callback cb, void* user) {
    if (!h) return PFB_ERR_ARG;
    if (n < 1 || !cb) PFB_FAIL(h, PFB_ERR_ARG, "host model needs n >= 1 and a callback");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    if (!h->copy_stream) PFB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {  // (the copy stream may already exist: pfb_elbo_batch creates it too)
        if (!h->hc_k[i]) PFB_CUDA(h, cudaEventCreateWithFlags(&h->hc_k[i], cudaEventDisableTiming));
        if (!h->hc_c[i]) PFB_CUDA(h, cudaEventCreateWithFlags(&h->hc_c[i], cudaEventDisableTiming));
    }
    h->host_cb = cb;
    h->host_user = user;
    h->model = PFB_MODEL_HOSTCALLBACK;
    h->model_n = n;
    return PFB_OK;
}

static int host_staging(pfb_engine* h, size_t x_bytes, size_t lp_bytes) {
    if (x_bytes > h->hX_cap) {
        for (int i = 0; i < 2; ++i) {
            if (h->hX[i]) cudaFreeHost(h->hX[i]);
            h->hX[i] = nullptr;
        }
        h->hX_cap = 0;
        for (int i = 0; i < 2; ++i) PFB_CUDA(h, cudaMallocHost((void**)&h->hX[i], x_bytes));
        h->hX_cap = x_bytes;
    }
    if (lp_bytes > h->hLp_cap) {
        if (h->hLp) cudaFreeHost(h->hLp);
        h->hLp = nullptr;
        h->hLp_cap = 0;
        PFB_CUDA(h, cudaMallocHost((void**)&h->hLp, lp_bytes));
        h->hLp_cap = lp_bytes;
    }
    return PFB_OK;
}

// dst[slot][k] = src[unit_of_slot][k] (NaN for unit < 0): log p of the best-iteration draws is the
// ELBO stage's own (K5 regenerates the same draws), so the host callback is not asked twice
__global__ void pfb_gather_unit_rows(int K, const int32_t* __restrict__ unit_of_slot, const double* __restrict__ src,
                                     double* __restrict__ dst) {
    const int u = unit_of_slot[blockIdx.x];
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        dst[(int64_t)blockIdx.x * K + k] = u >= 0 ? src[(int64_t)u * K + k] : NAN;
}

// log p of M = cnt * K materialised device draws through the host callback (synchronous).
static int host_logp_sync(pfb_engine* h, const double* dX, int64_t M, double* d_logp) {
    const size_t n = (size_t)h->n;
    int rc = host_staging(h, (size_t)M * n * 8, (size_t)M * 8);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    PFB_CUDA(h, cudaMemcpyAsync(h->hX[0], dX, (size_t)M * n * 8, cudaMemcpyDeviceToHost, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));
    h->host_cb(h->host_user, h->hX[0], (int64_t)n, M, h->hLp);
    PFB_CUDA(h, cudaMemcpyAsync(d_logp, h->hLp, (size_t)M * 8, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));
    return PFB_OK;
}

// K8g: log p of M = nslots * K materialised draws X [n x M] for the GEMM-shaped families — one fused
// FP64 tensor-core GEMM + reduction kernel (k8_gemm_logp.cu); no library GEMM, no product matrix in HBM.
static int generic_logp(pfb_engine* h, const double* X, int64_t M, const int32_t* slot_unit, double* logp,
                        int K_over = 0) {
    if (M <= 0) return PFB_OK;
    const int n = h->n, K = K_over > 0 ? K_over : h->K;
    if (h->model == PFB_MODEL_DENSENORMAL) {
        const double* d = h->dModel.as<double>();  // { m[n], P m[n], P[n x n] }
        PFB_CUDA(h, pfb_launch_k8g_dense(h->stream, n, M, K, slot_unit, X, d + 2 * (size_t)n, d + n, h->model_c0, logp));
    } else {
        const int nobs = h->model_nobs, nb = n - 2;
        const double* Xm = h->dModel.as<double>();  // { X[nobs x (n-2)], y[nobs], X' }
        const double* yobs = Xm + (size_t)nobs * nb;
        PFB_CUDA(h, pfb_launch_k8g_logistic(h->stream, n, nobs, M, K, slot_unit, X, Xm, yobs, logp));
    }
    return PFB_OK;
}

// K_over > 0 / seeds_over != NULL: fresh draws from the fitted normals (top-up draws, resample()).
static cudaError_t launch_k3(pfb_engine* h, int nslots, const int32_t* unit_list, double* logp, double* logq,
                             double* draws, int K_over, const uint64_t* seeds_over, const int32_t* sel_cnt,
                             const void* sel_list, int sel_cap, const uint64_t* fb_seeds, const int32_t* fb_paths) {
    const double* mp0 = model_is_external(h) ? nullptr : h->dModel.as<double>();
    const double* mp1 = mp0 ? mp0 + h->model_n : nullptr;
    const double* un = (h->have_normals && (!seeds_over || seeds_over == h->dSeeds.as<uint64_t>()))
                           ? h->dNormals.as<double>() : nullptr;
    if (h->generic) {
        cudaError_t e = h->dKgScratch.ensure(pfb_kg_k3_scratch_doubles(h->KP, h->n, nslots) * 8 + 16);
        if (e != cudaSuccess) return e;
        return pfb_launch_k3g(h->stream, h->KP, h->model, h->n, K_over > 0 ? K_over : h->K, nslots, unit_list,
                              h->dFR.as<double>(), h->dHDR.as<double>(), seeds_over ? seeds_over : h->dSeeds.as<uint64_t>(),
                              un, mp0, mp1, h->model_c0, logp, logq, draws, sel_cnt, sel_list, sel_cap,
                              fb_seeds ? h->dX.as<double>() : nullptr, h->dG.as<double>(), h->dOff.as<int64_t>(), fb_seeds,
                              fb_paths, h->dKgScratch.as<double>());
    }
    auto fn = h->KP == 12 ? pfb_launch_k3_kp12 : (h->KP == 20 ? pfb_launch_k3_kp20 : pfb_launch_k3_kp24);
    return fn(h->stream, h->model, h->n, K_over > 0 ? K_over : h->K, nslots, unit_list, h->dFR2.as<double>(),
              h->dHDR.as<double>(), seeds_over ? seeds_over : h->dSeeds.as<uint64_t>(), un, mp0, mp1, h->model_c0,
              logp, logq, draws, h->cfg.elbo_mode == 1, sel_cnt, sel_list, sel_cap,
              fb_seeds ? h->dX.as<double>() : nullptr, h->dG.as<double>(), h->dOff.as<int64_t>(), fb_seeds, fb_paths);
}

// K5, on demand: materialise the best-iteration draws of every path into the pool (regenerated, hence
// identical to the ELBO-stage draws because the RNG is counter based).  The pool's logp / logq are
// gathered from the ELBO stage by pfb_batch_run; the draws are only produced when somebody looks at
// them (PathfinderResult.draws, pool exchange) — the resample stage regenerates just its columns.
static int resolve_failed(pfb_engine* h);
static int ensure_pool(pfb_engine* h) {
    if (h->pool_ready || h->P <= 0) return PFB_OK;
    if (!h->ran || h->poolK != h->K || h->poolP != h->P)
        PFB_FAIL(h, PFB_ERR_STATE, "no device pool (pfb_batch_run / pfb_draw_from_fits)");
    PFB_CUDA(h, launch_k3(h, h->P, h->dBestUnit.as<int32_t>(), nullptr, nullptr, h->dPool.as<double>(), 0,
                          h->pool_seeds, nullptr, nullptr, 0, h->dFbSeeds.as<uint64_t>()));
    h->launches += 1;
    h->pool_ready = true;
    return PFB_OK;
}

// Columns inds[t] (1-based pool indices; this engine's pool covers [base, base + P K)) regenerated
// straight into d_out [n x m]; columns owned elsewhere are left untouched.
static int regen_columns(pfb_engine* h, int m, const int64_t* d_inds, int64_t base, double* d_out) {
    if (m <= 0 || h->P <= 0) return PFB_OK;
    cudaStream_t st = h->stream;
    PFB_CUDA(h, h->dSelCnt.ensure((size_t)h->P * 4 + 8));
    PFB_CUDA(h, h->dSelList.ensure((size_t)h->P * (size_t)m * 8 + 8));
    PFB_CUDA(h, pfb_launch_k7r_bin(st, h->P, h->K, m, base, d_inds, h->dSelCnt.as<int32_t>(), h->dSelList.p));
    PFB_CUDA(h, launch_k3(h, h->P, h->dBestUnit.as<int32_t>(), nullptr, nullptr, d_out, 0, h->pool_seeds,
                          h->dSelCnt.as<int32_t>(), h->dSelList.p, m, h->dFbSeeds.as<uint64_t>()));
    h->launches += 2;
    return PFB_OK;
}

// ELBO stage for a host-callback target (row f2): K3 materialises chunk c + 2 and the copy stream
// brings chunk c + 1 to pinned memory while the host evaluates chunk c.
static int run_host_callback_stage(pfb_engine* h) {
    const int n = h->n, K = h->K, U = (int)h->U;
    const size_t per_unit = (size_t)n * (size_t)K * 8;
    int chunk = (int)std::max<size_t>(1, ((size_t)64 << 20) / per_unit);
    chunk = std::min(chunk, U);
    const bool mat = h->cfg.materialize_all != 0;
    cudaStream_t st = h->stream, cs = h->copy_stream;
    if (!mat) PFB_CUDA(h, h->dGenX.ensure(2 * per_unit * (size_t)chunk));
    PFB_CUDA(h, h->dIota.ensure((size_t)U * 4));
    {
        std::vector<int32_t> iota((size_t)U);
        for (int u = 0; u < U; ++u) iota[(size_t)u] = u;
        PFB_CUDA(h, cudaMemcpyAsync(h->dIota.p, iota.data(), (size_t)U * 4, cudaMemcpyHostToDevice, st));
        PFB_CUDA(h, cudaStreamSynchronize(st));
    }
    int rc = host_staging(h, per_unit * (size_t)chunk, (size_t)U * K * 8);
    if (rc) return rc;
    const int nchunks = (U + chunk - 1) / chunk;
    h->host_cb_s = 0.0;
    auto sample_and_copy = [&](int c) -> int {
        const int u0 = c * chunk, cnt = std::min(chunk, U - u0), b = c & 1;
        double* xbuf = mat ? h->dAllDraws.as<double>() + (size_t)u0 * n * K
                           : h->dGenX.as<double>() + (size_t)b * chunk * n * K;
        if (c >= 2) PFB_CUDA(h, cudaStreamWaitEvent(st, h->hc_c[b], 0));  // device buffer b drained
        PFB_CUDA(h, launch_k3(h, cnt, h->dIota.as<int32_t>() + u0, h->dLogp.as<double>() + (size_t)u0 * K,
                              h->dLogq.as<double>() + (size_t)u0 * K, xbuf, 0, nullptr));
        h->launches += 1;
        PFB_CUDA(h, cudaEventRecord(h->hc_k[b], st));
        PFB_CUDA(h, cudaStreamWaitEvent(cs, h->hc_k[b], 0));
        PFB_CUDA(h, cudaMemcpyAsync(h->hX[b], xbuf, per_unit * (size_t)cnt, cudaMemcpyDeviceToHost, cs));
        PFB_CUDA(h, cudaEventRecord(h->hc_c[b], cs));
        return PFB_OK;
    };
    for (int c = 0; c < std::min(2, nchunks); ++c) {
        rc = sample_and_copy(c);
        if (rc) return rc;
    }
    for (int c = 0; c < nchunks; ++c) {
        const int u0 = c * chunk, cnt = std::min(chunk, U - u0), b = c & 1;
        PFB_CUDA(h, cudaEventSynchronize(h->hc_c[b]));
        double* lp = h->hLp + (size_t)u0 * K;
        const auto t0 = std::chrono::steady_clock::now();
        h->host_cb(h->host_user, h->hX[b], (int64_t)n, (int64_t)cnt * K, lp);
        h->host_cb_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        PFB_CUDA(h, cudaMemcpyAsync(h->dLogp.as<double>() + (size_t)u0 * K, lp, (size_t)cnt * K * 8,
                                    cudaMemcpyHostToDevice, st));
        if (c + 2 < nchunks) {
            rc = sample_and_copy(c + 2);
            if (rc) return rc;
        }
    }
    return PFB_OK;
}

__global__ void pfb_iota(int n, int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
__global__ void pfb_set_pool_seeds(int nf, const int64_t* __restrict__ pairs, uint64_t* __restrict__ pool_seeds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // pairs: (unit, seed)
    if (i < nf && pairs[2 * i] >= 0) pool_seeds[pairs[2 * i]] = (uint64_t)pairs[2 * i + 1];
}
__global__ void pfb_scatter_rows(int K, const int32_t* __restrict__ slot_of_row, const double* __restrict__ src,
                                 double* __restrict__ dst) {
    const int s = slot_of_row[blockIdx.x];
    for (int k = threadIdx.x; k < K; k += blockDim.x) dst[(int64_t)s * K + k] = src[(int64_t)blockIdx.x * K + k];
}

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// Failed paths (success = false: no iteration, or a NaN / -Inf best ELBO).  The reference draws
//   rand(rng, fit_distributions[fit_iteration + 1], ndraws)       (src/singlepath.jl:224-228)
// for them — fresh draws with the path's rng, which then enter the PSIS pool with their own log
// densities (src/multipath.jl:217, src/resample.jl:81-95).  Here: K fresh draws per failed path from
// the fit of its best iteration (the identity fit of iteration 0 when it has none) with the path's
// fallback seed; the pool's log densities and the pool's seed table are updated, so that a later
// materialisation / column regeneration reproduces exactly these draws.
static int failed_path_enqueue(pfb_engine* h) {
    const int P = h->P;
    cudaStream_t st = h->stream;
    h->pool_seeds = h->dSeeds.as<uint64_t>();
    h->n_failed = 0;
    h->fail_pending = false;
    if (P <= 0) return PFB_OK;
    h->fb_seeds.resize((size_t)P);
    for (int p = 0; p < P; ++p)
        h->fb_seeds[(size_t)p] = ((int)h->fb_seeds_user.size() == P) ? h->fb_seeds_user[(size_t)p]
                                                                      : splitmix64(0x5EEDFA11ULL + (uint64_t)p);
    h->fb_seeds_user.clear();
    PFB_CUDA(h, h->dFbSeeds.ensure((size_t)P * 8));
    PFB_CUDA(h, cudaMemcpyAsync(h->dFbSeeds.p, h->fb_seeds.data(), (size_t)P * 8, cudaMemcpyHostToDevice, st));
    if (P > h->hSuccBuCap) {
        PFB_CUDA(h, cudaStreamSynchronize(st));  // an earlier batch's flags may still be on their way into the old pair
        if (h->hSuccBu) cudaFreeHost(h->hSuccBu);
        h->hSuccBu = nullptr;
        h->hSuccBuCap = 0;
        PFB_CUDA(h, cudaMallocHost((void**)&h->hSuccBu, (size_t)P * 8));
        h->hSuccBuCap = P;
    }
    if (!h->ev_fail) PFB_CUDA(h, cudaEventCreateWithFlags(&h->ev_fail, cudaEventDisableTiming));
    PFB_CUDA(h, cudaMemcpyAsync(h->hSuccBu, h->dSucc.p, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    PFB_CUDA(h, cudaMemcpyAsync(h->hSuccBu + h->hSuccBuCap, h->dBestUnit.p, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    PFB_CUDA(h, cudaEventRecord(h->ev_fail, st));
    h->fail_pending = true;
    return PFB_OK;
}

// Second half, run by the first consumer of the pool (or by the optimistic consumers after their own
// synchronisation, see pfb_psis_resample / pfb_batch_download): waits for the flags, and for failed
// paths replaces their pool entries as described above.  No failed path (the common case): nothing to do.
static int resolve_failed(pfb_engine* h) {
    if (!h->fail_pending) return PFB_OK;
    h->fail_pending = false;
    const int P = h->P, K = h->K, n = h->n;
    cudaStream_t st = h->stream;
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    PFB_CUDA(h, cudaEventSynchronize(h->ev_fail));
    const int32_t* succ = h->hSuccBu;
    const int32_t* bu = h->hSuccBu + h->hSuccBuCap;
    std::vector<int32_t> paths, units;
    std::vector<int64_t> pairs;
    for (int p = 0; p < P; ++p)
        if (!succ[(size_t)p]) {
            paths.push_back(p);
            units.push_back(bu[(size_t)p]);
            pairs.push_back(bu[(size_t)p]);
            pairs.push_back((int64_t)h->fb_seeds[(size_t)p]);
        }
    const int nf = (int)paths.size();
    h->n_failed = nf;
    if (nf == 0 || h->have_normals) return PFB_OK;  // (parity mode supplies the normals of the ELBO stage only)
    h->pool_ready = false;  // an optimistically materialised pool holds the wrong draws for the failed paths
    const int64_t U = h->U;
    PFB_CUDA(h, h->dPoolSeeds.ensure((size_t)std::max<int64_t>(U, 1) * 8));
    PFB_CUDA(h, h->dFbUnits.ensure((size_t)nf * 4));
    PFB_CUDA(h, h->dFbPaths.ensure((size_t)nf * 4));
    PFB_CUDA(h, h->dFbPairs.ensure((size_t)nf * 16));
    PFB_CUDA(h, h->dFbLogp.ensure((size_t)nf * K * 8));
    PFB_CUDA(h, h->dFbLogq.ensure((size_t)nf * K * 8));
    if (U > 0)
        PFB_CUDA(h, cudaMemcpyAsync(h->dPoolSeeds.p, h->dSeeds.p, (size_t)U * 8, cudaMemcpyDeviceToDevice, st));
    PFB_CUDA(h, cudaMemcpyAsync(h->dFbUnits.p, units.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, cudaMemcpyAsync(h->dFbPaths.p, paths.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, cudaMemcpyAsync(h->dFbPairs.p, pairs.data(), (size_t)nf * 16, cudaMemcpyHostToDevice, st));
    pfb_set_pool_seeds<<<(nf + 127) / 128, 128, 0, st>>>(nf, h->dFbPairs.as<int64_t>(), h->dPoolSeeds.as<uint64_t>());
    PFB_CUDA(h, cudaGetLastError());
    h->pool_seeds = h->dPoolSeeds.as<uint64_t>();
    const bool ext = h->model == PFB_MODEL_DENSENORMAL || h->model == PFB_MODEL_HLOGISTIC ||
                     h->model == PFB_MODEL_HOSTCALLBACK;
    double* xbuf = nullptr;
    if (ext) {
        PFB_CUDA(h, h->dGenX.ensure((size_t)nf * (size_t)n * (size_t)K * 8 + 8));
        xbuf = h->dGenX.as<double>();
    }
    PFB_CUDA(h, launch_k3(h, nf, h->dFbUnits.as<int32_t>(), h->dFbLogp.as<double>(), h->dFbLogq.as<double>(), xbuf, 0,
                          h->pool_seeds, nullptr, nullptr, 0, h->dFbSeeds.as<uint64_t>(), h->dFbPaths.as<int32_t>()));
    if (h->model == PFB_MODEL_HOSTCALLBACK) {
        int rc = host_logp_sync(h, xbuf, (int64_t)nf * K, h->dFbLogp.as<double>());
        if (rc) return rc;
    } else if (ext) {
        int rc = generic_logp(h, xbuf, (int64_t)nf * K, nullptr, h->dFbLogp.as<double>(), 0);
        if (rc) return rc;
    }
    pfb_scatter_rows<<<nf, 256, 0, st>>>(K, h->dFbPaths.as<int32_t>(), h->dFbLogp.as<double>(), h->dPoolLogp.as<double>());
    pfb_scatter_rows<<<nf, 256, 0, st>>>(K, h->dFbPaths.as<int32_t>(), h->dFbLogq.as<double>(), h->dPoolLogq.as<double>());
    PFB_CUDA(h, cudaGetLastError());
    h->launches += 4;
    return PFB_OK;
}

extern "C" int pfb_set_fallback_seeds(pfb_handle h, int P, const uint64_t* seeds) {
    if (!h || P < 0 || (P > 0 && !seeds)) return PFB_ERR_ARG;
    h->fb_seeds_user.assign(seeds, seeds + P);
    return PFB_OK;
}

extern "C" int pfb_batch_run(pfb_handle h) {
    if (!h) return PFB_ERR_ARG;
    if (!h->have_batch) PFB_FAIL(h, PFB_ERR_STATE, "pfb_batch_upload has not been called");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const int n = h->n, P = h->P, K = h->K, J = h->cfg.history_length, KP = h->KP;
    const int U = (int)h->U;
    h->launches = 0;
    PFB_CUDA(h, cudaEventRecord(h->ev[0], st));
    const int k2_model = model_is_external(h) ? PFB_MODEL_ISONORMAL : h->model;
    const double* k2_mp0 = model_is_external(h) ? nullptr : h->dModel.as<double>();
    const double* k2_mp1 = (h->dModel.p && !model_is_external(h)) ? h->dModel.as<double>() + h->model_n : nullptr;
    const bool k3_by_group = h->up_ngroups >= 2 && !model_is_external(h) && U > 0;
    int k3_done = 0;
    if (k3_by_group) {
        PFB_CUDA(h, h->dIota.ensure((size_t)U * 4));
        pfb_iota<<<(U + 255) / 256, 256, 0, st>>>(U, h->dIota.as<int32_t>());
        PFB_CUDA(h, cudaGetLastError());
    }
    if (h->up_ngroups > 0) {
        // pfb_elbo_batch: the trajectories are still arriving group by group on the copy stream; K1 and
        // K2 of a group start as soon as its points are resident (timers: k1 = first group, k2 = the rest)
        for (int g = 0; g < h->up_ngroups; ++g) {
            const int p0 = h->up_p[g], p1 = h->up_p[g + 1];
            const int u0 = (int)(h->h_off[(size_t)p0] - p0), u1 = (int)(h->h_off[(size_t)p1] - p1);
            // K1 takes the same time for 16 paths as for 64 (a sequential chain per path), so its
            // groups run on their own stream, beside the copies and beside K2 of the previous group
            PFB_CUDA(h, cudaStreamWaitEvent(h->k1_stream, h->up_ev[g], 0));
            PFB_CUDA(h, pfb_launch_k1_range(h->k1_stream, n, p0, p1 - p0, J, h->cfg.eps, h->dX.as<double>(),
                                            h->dG.as<double>(), h->dOff.as<int64_t>(), h->dAlpha.as<double>(),
                                            h->dHist.as<int32_t>(), h->dHistCnt.as<int32_t>(), h->dRej.as<int64_t>()));
            PFB_CUDA(h, cudaEventRecord(h->k1_ev[g], h->k1_stream));
            PFB_CUDA(h, cudaStreamWaitEvent(st, h->k1_ev[g], 0));
            if (g == 0) PFB_CUDA(h, cudaEventRecord(h->ev[1], st));
            PFB_CUDA(h, launch_k2_any(h, u0, u1 - u0, k2_model, k2_mp0, k2_mp1));
            h->launches += (p1 > p0) + (u1 > u0);
            if (k3_by_group && (g == h->up_ngroups / 2 - 1 || g == h->up_ngroups - 1)) {
                // K3 in two launches, after the first and the second half of the groups: the copy stream
                // keeps uploading the second half while the tensor cores already work on the first.  Measured
                // (config 3, one B200, end-to-end step): one launch 14.76 ms, two halves 14.70 ms, a quarter
                // then the rest 15.05 ms, one launch per group 15.18 ms (every launch adds a partial last wave)
                const int ua = (g == h->up_ngroups - 1) ? k3_done : 0;
                if (ua == 0) PFB_CUDA(h, cudaEventRecord(h->ev[2], st));
                if (u1 > ua) {
                    PFB_CUDA(h, launch_k3(h, u1 - ua, h->dIota.as<int32_t>() + ua, h->dLogp.as<double>() + (size_t)ua * K,
                                          h->dLogq.as<double>() + (size_t)ua * K,
                                          h->cfg.materialize_all ? h->dAllDraws.as<double>() + (size_t)ua * n * K : nullptr));
                    h->launches += 1;
                }
                k3_done = u1;
            }
        }
        h->up_ngroups = 0;
    } else {
        PFB_CUDA(h, pfb_launch_k1(st, n, P, J, h->cfg.eps, h->dX.as<double>(), h->dG.as<double>(),
                                  h->dOff.as<int64_t>(), h->dAlpha.as<double>(), h->dHist.as<int32_t>(),
                                  h->dHistCnt.as<int32_t>(), h->dRej.as<int64_t>()));
        h->launches += (P > 0);
        PFB_CUDA(h, cudaEventRecord(h->ev[1], st));
        PFB_CUDA(h, launch_k2_any(h, 0, U, k2_model, k2_mp0, k2_mp1));
        h->launches += (U > 0);
    }
    if (!k3_by_group) PFB_CUDA(h, cudaEventRecord(h->ev[2], st));
    if (k3_by_group) {
        // (already launched group by group)
    } else if (!model_is_external(h)) {
        PFB_CUDA(h, launch_k3(h, U, nullptr, h->dLogp.as<double>(), h->dLogq.as<double>(),
                              h->cfg.materialize_all ? h->dAllDraws.as<double>() : nullptr));
        h->launches += (U > 0);
    } else if (h->model == PFB_MODEL_HOSTCALLBACK) {
        if (U > 0) {
            int rc = run_host_callback_stage(h);
            if (rc) return rc;
        }
    } else if (U > 0) {
        // GEMM-shaped log p: materialise the draws of a chunk of units (K3), then K8g (fused FP64 tensor-core GEMM + log p reduction)
        const size_t per_unit = (size_t)n * (size_t)K * 8;
        int chunk = (int)std::max<size_t>(1, ((size_t)4 << 30) / per_unit);
        if (h->cfg.materialize_all || chunk > U) chunk = U;
        if (chunk < U) {
            // K8g runs one CTA per 128 draws: prefer a chunk whose CTA count fills its last wave of SMs
            int nsm = 148;
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->cfg.device);
            int best = chunk;
            double best_fill = 0.0;
            for (int c = chunk; c >= std::max(1, chunk / 2); --c) {
                const int64_t ctas = ((int64_t)c * K + 127) / 128;
                const double fill = (double)ctas / (double)(((ctas + nsm - 1) / nsm) * nsm);
                if (fill > best_fill + 1e-9) { best_fill = fill; best = c; }
            }
            chunk = best;
        }
        if (!h->cfg.materialize_all) PFB_CUDA(h, h->dGenX.ensure(per_unit * (size_t)chunk));
        PFB_CUDA(h, h->dIota.ensure((size_t)U * 4));
        {
            std::vector<int32_t> iota((size_t)U);
            for (int u = 0; u < U; ++u) iota[(size_t)u] = u;
            PFB_CUDA(h, cudaMemcpyAsync(h->dIota.p, iota.data(), (size_t)U * 4, cudaMemcpyHostToDevice, st));
            PFB_CUDA(h, cudaStreamSynchronize(st));
        }
        for (int u0 = 0; u0 < U; u0 += chunk) {
            const int cnt = std::min(chunk, U - u0);
            double* xbuf = h->cfg.materialize_all ? h->dAllDraws.as<double>() + (size_t)u0 * n * K
                                                  : h->dGenX.as<double>();
            double* lp = h->dLogp.as<double>() + (size_t)u0 * K;
            PFB_CUDA(h, launch_k3(h, cnt, h->dIota.as<int32_t>() + u0, lp, h->dLogq.as<double>() + (size_t)u0 * K,
                                  xbuf));
            int rc = generic_logp(h, xbuf, (int64_t)cnt * K, nullptr, lp);
            if (rc) return rc;
            h->launches += 2;
        }
    }
    PFB_CUDA(h, cudaEventRecord(h->ev[3], st));
    PFB_CUDA(h, pfb_launch_k4(st, P, K, (int64_t)U, h->dOff.as<int64_t>(), h->dLogp.as<double>(), h->dLogq.as<double>(),
                              h->dElbo.as<double>(), h->dSe.as<double>(), h->dBestIter.as<int64_t>(),
                              h->dBestUnit.as<int32_t>(), h->dSucc.as<int32_t>()));
    h->launches += (P > 0) + (U > 0);
    PFB_CUDA(h, cudaEventRecord(h->ev[4], st));
    // the pool's log densities are the ELBO stage's own (same draws); its draws are materialised on
    // demand (ensure_pool) or column by column (regen_columns)
    if (P > 0) {
        pfb_gather_unit_rows<<<P, 256, 0, st>>>(K, h->dBestUnit.as<int32_t>(), h->dLogp.as<double>(),
                                                h->dPoolLogp.as<double>());
        pfb_gather_unit_rows<<<P, 256, 0, st>>>(K, h->dBestUnit.as<int32_t>(), h->dLogq.as<double>(),
                                                h->dPoolLogq.as<double>());
        PFB_CUDA(h, cudaGetLastError());
        h->launches += 2;
    }
    h->pool_ready = false;
    h->ran = true;
    h->poolK = K;
    h->poolP = P;
    {
        int rcf = failed_path_enqueue(h);
        if (rcf) return rcf;
    }
    PFB_CUDA(h, cudaEventRecord(h->ev[5], st));
    return PFB_OK;
}

// K1 + K2 only, with the best iteration of every path given by the caller: rebuilds the fitted
// normals of a stored result (resample() re-entry, src/resample.jl:20-46) without an ELBO stage.
extern "C" int pfb_batch_fit_only(pfb_handle h, const int64_t* best_iter) {
    if (!h || !best_iter) return PFB_ERR_ARG;
    if (!h->have_batch) PFB_FAIL(h, PFB_ERR_STATE, "pfb_batch_upload has not been called");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const int n = h->n, P = h->P, J = h->cfg.history_length, KP = h->KP;
    const int U = (int)h->U;
    std::vector<int32_t> bu((size_t)P);
    std::vector<int32_t> ok((size_t)P);
    for (int p = 0; p < P; ++p) {
        const int64_t L = h->h_off[(size_t)p + 1] - h->h_off[(size_t)p] - 1;
        if (best_iter[p] < 0 || best_iter[p] > L) PFB_FAIL(h, PFB_ERR_ARG, "best_iter out of range");
        bu[(size_t)p] = best_iter[p] > 0 ? (int32_t)(h->h_off[(size_t)p] - p + best_iter[p] - 1) : -1;
        ok[(size_t)p] = best_iter[p] > 0;
    }
    PFB_CUDA(h, pfb_launch_k1(st, n, P, J, h->cfg.eps, h->dX.as<double>(), h->dG.as<double>(),
                              h->dOff.as<int64_t>(), h->dAlpha.as<double>(), h->dHist.as<int32_t>(),
                              h->dHistCnt.as<int32_t>(), h->dRej.as<int64_t>()));
    PFB_CUDA(h, launch_k2_any(h, 0, U, model_is_external(h) ? PFB_MODEL_ISONORMAL : h->model,
                              model_is_external(h) ? nullptr : h->dModel.as<double>(),
                              (h->dModel.p && !model_is_external(h)) ? h->dModel.as<double>() + h->model_n : nullptr));
    PFB_CUDA(h, cudaMemcpyAsync(h->dBestUnit.p, bu.data(), (size_t)P * 4, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, cudaMemcpyAsync(h->dBestIter.p, best_iter, (size_t)P * 8, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, cudaMemcpyAsync(h->dSucc.p, ok.data(), (size_t)P * 4, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));
    for (int i = 0; i <= 5; ++i) PFB_CUDA(h, cudaEventRecord(h->ev[i], st));
    h->launches = 2;
    h->ran = true;
    h->pool_seeds = h->dSeeds.as<uint64_t>();
    h->n_failed = 0;
    h->fail_pending = false;
    h->poolK = 0;
    h->poolP = 0;
    h->pool_ready = false;
    return PFB_OK;
}

// K_new fresh draws from the best-iteration normal of every path, seeded per path:
//   rand(rng, fit_distribution, K_new)  of src/singlepath.jl:228-230 (top-up draws) and
//   src/resample.jl:102-109 (resample with ndraws_per_run), with their logp and logq = logpdf(fit, x)
//   (what _compute_log_importance_ratios, src/resample.jl:81-95, evaluates for fresh draws).
// keep_as_pool != 0: the device pool becomes these draws (N = P * K_new for pfb_psis_resample).
extern "C" int pfb_draw_from_fits(pfb_handle h, int K_new, const uint64_t* seeds, double* draws, double* logp,
                                  double* logq, int keep_as_pool) {
    if (!h || !seeds) return PFB_ERR_ARG;
    if (K_new < 1) PFB_FAIL(h, PFB_ERR_ARG, "K_new must be positive");
    if (!h->ran) PFB_FAIL(h, PFB_ERR_STATE, "no fitted batch (pfb_batch_run / pfb_batch_fit_only)");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    { int rf = resolve_failed(h); if (rf) return rf; }
    cudaStream_t st = h->stream;
    const size_t n = h->n, P = h->P, U = (size_t)h->U;
    if (P == 0) return PFB_OK;
    // K3 reads seeds[unit]: scatter the per-path seeds to their best units
    std::vector<int32_t> bu(P);
    PFB_CUDA(h, cudaMemcpyAsync(bu.data(), h->dBestUnit.p, P * 4, cudaMemcpyDeviceToHost, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));
    std::vector<uint64_t> sd(U > 0 ? U : 1, 0);
    for (size_t p = 0; p < P; ++p)
        if (bu[p] >= 0) sd[(size_t)bu[p]] = seeds[p];
    PFB_CUDA(h, h->dTopSeeds.ensure(sd.size() * 8));
    PFB_CUDA(h, cudaMemcpyAsync(h->dTopSeeds.p, sd.data(), sd.size() * 8, cudaMemcpyHostToDevice, st));
    DevBuf* X = keep_as_pool ? &h->dPool : &h->dGenX;
    DevBuf* Lp = keep_as_pool ? &h->dPoolLogp : &h->dTmpLogr;
    DevBuf* Lq = keep_as_pool ? &h->dPoolLogq : &h->dTmpPool;
    PFB_CUDA(h, X->ensure(n * (size_t)K_new * P * 8 + 8));
    PFB_CUDA(h, Lp->ensure((size_t)K_new * P * 8 + 8));
    PFB_CUDA(h, Lq->ensure((size_t)K_new * P * 8 + 8));
    PFB_CUDA(h, h->dTopFb.ensure(P * 8));
    PFB_CUDA(h, cudaMemcpyAsync(h->dTopFb.p, seeds, P * 8, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, launch_k3(h, (int)P, h->dBestUnit.as<int32_t>(), Lp->as<double>(), Lq->as<double>(), X->as<double>(),
                          K_new, h->dTopSeeds.as<uint64_t>(), nullptr, nullptr, 0, h->dTopFb.as<uint64_t>()));
    if (h->model == PFB_MODEL_HOSTCALLBACK) {
        int rc = host_logp_sync(h, X->as<double>(), (int64_t)P * K_new, Lp->as<double>());
        if (rc) return rc;
    } else if (model_is_external(h)) {
        int rc = generic_logp(h, X->as<double>(), (int64_t)P * K_new, h->dBestUnit.as<int32_t>(), Lp->as<double>(),
                              K_new);
        if (rc) return rc;
    }
    if (draws) PFB_CUDA(h, cudaMemcpyAsync(draws, X->p, n * (size_t)K_new * P * 8, cudaMemcpyDeviceToHost, st));
    if (logp) PFB_CUDA(h, cudaMemcpyAsync(logp, Lp->p, (size_t)K_new * P * 8, cudaMemcpyDeviceToHost, st));
    if (logq) PFB_CUDA(h, cudaMemcpyAsync(logq, Lq->p, (size_t)K_new * P * 8, cudaMemcpyDeviceToHost, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));
    if (keep_as_pool) {
        h->poolK = K_new;
        h->poolP = h->P;
        h->pool_ready = true;
    }
    return PFB_OK;
}

// Draws, log p and log q of arbitrary (path, iteration) units of the CURRENT batch, regenerated
// from their seeds (counter-based RNG => identical to what the ELBO stage saw): the
// ELBOEstimate.draws / .log_densities_* payload of src/elbo.jl:22-29 on demand, instead of keeping
// n x K doubles for every iteration like the reference does.
extern "C" int pfb_unit_draws(pfb_handle h, int nunits, const int32_t* units, double* draws, double* logp,
                              double* logq) {
    if (!h || (nunits > 0 && !units)) return PFB_ERR_ARG;
    if (!h->ran) PFB_FAIL(h, PFB_ERR_STATE, "no fitted batch (pfb_batch_run)");
    if (nunits <= 0) return PFB_OK;
    for (int i = 0; i < nunits; ++i)
        if (units[i] < 0 || units[i] >= h->U) PFB_FAIL(h, PFB_ERR_ARG, "unit index out of range");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const size_t n = h->n, K = h->K, M = (size_t)nunits * K;
    PFB_CUDA(h, h->dIota.ensure((size_t)std::max<int64_t>(h->U, nunits) * 4 + (size_t)nunits * 4));
    // the unit list lives behind the iota area so that a running external-model batch is not disturbed
    int32_t* d_units = h->dIota.as<int32_t>() + std::max<int64_t>(h->U, nunits);
    PFB_CUDA(h, cudaMemcpyAsync(d_units, units, (size_t)nunits * 4, cudaMemcpyHostToDevice, st));
    PFB_CUDA(h, h->dGenX.ensure(M * n * 8 + 8));
    PFB_CUDA(h, h->dTmpLogr.ensure(M * 8 + 8));
    PFB_CUDA(h, h->dTmpPool.ensure(M * 8 + 8));
    double* X = h->dGenX.as<double>();
    double* Lp = h->dTmpLogr.as<double>();
    double* Lq = h->dTmpPool.as<double>();
    PFB_CUDA(h, launch_k3(h, nunits, d_units, Lp, Lq, X));
    if (h->model == PFB_MODEL_HOSTCALLBACK) {
        int rc = host_logp_sync(h, X, (int64_t)M, Lp);
        if (rc) return rc;
    } else if (model_is_external(h)) {
        int rc = generic_logp(h, X, (int64_t)M, d_units, Lp);
        if (rc) return rc;
    }
    if (draws) PFB_CUDA(h, cudaMemcpyAsync(draws, X, M * n * 8, cudaMemcpyDeviceToHost, st));
    if (logp) PFB_CUDA(h, cudaMemcpyAsync(logp, Lp, M * 8, cudaMemcpyDeviceToHost, st));
    if (logq) PFB_CUDA(h, cudaMemcpyAsync(logq, Lq, M * 8, cudaMemcpyDeviceToHost, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));
    return PFB_OK;
}

static int gather_fits(pfb_engine* h, int cnt, const int32_t* d_units, double* mu, double* alpha, double* vh,
                       double* T, double* Vc, double* logdet, int32_t* jeff);

// fit_distributions[l + 1] of arbitrary (path, iteration) units of the current batch in the reference's
// WoodburyPDMat form (src/singlepath.jl:64 keeps every iteration's; here they are exported on demand).
extern "C" int pfb_unit_fits(pfb_handle h, int nunits, const int32_t* units, double* mu, double* alpha, double* vh,
                             double* T, double* Vc, double* logdet, int32_t* jeff) {
    if (!h || (nunits > 0 && !units)) return PFB_ERR_ARG;
    if (!h->ran) PFB_FAIL(h, PFB_ERR_STATE, "no fitted batch (pfb_batch_run / pfb_batch_fit_only)");
    if (nunits <= 0) return PFB_OK;
    for (int i = 0; i < nunits; ++i)
        if (units[i] < 0 || units[i] >= h->U) PFB_FAIL(h, PFB_ERR_ARG, "unit index out of range");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    PFB_CUDA(h, h->dIota.ensure((size_t)std::max<int64_t>(h->U, nunits) * 4 + (size_t)nunits * 4));
    int32_t* d_units = h->dIota.as<int32_t>() + std::max<int64_t>(h->U, nunits);
    PFB_CUDA(h, cudaMemcpyAsync(d_units, units, (size_t)nunits * 4, cudaMemcpyHostToDevice, h->stream));
    return gather_fits(h, nunits, d_units, mu, alpha, vh, T, Vc, logdet, jeff);
}

extern "C" int pfb_batch_sync(pfb_handle h) {
    if (!h) return PFB_ERR_ARG;
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    PFB_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFB_OK;
}

// gather of the best-iteration factor in the reference's WoodburyPDMat form (from the swizzled
// FR2 record, pfb_common.cuh)
__global__ void pfb_gather_fit(int n, int KP, const int32_t* __restrict__ best_unit,
                               const double* __restrict__ FR2, const double* __restrict__ HDR,
                               const double* __restrict__ alpha, const int32_t* __restrict__ hist_cnt,
                               double* mu, double* al, double* vh, double* Tm, double* Vc, double* logdet,
                               int32_t* jeff) {
    const int p = blockIdx.x;
    const int u = best_unit[p];
    const int RS2 = pfb_rs2_of(KP), HS = pfb_hs_of(KP), npad = pfb_npad8(n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (u >= 0) {
            const double* row = FR2 + ((int64_t)u * npad + i) * RS2;
            const int sw = pfb_swz(i);
            mu[(int64_t)p * n + i] = row[(KP + 1) ^ sw];
            al[(int64_t)p * n + i] = alpha[(int64_t)u * n + i];
            for (int j = 0; j < KP; ++j) vh[((int64_t)p * KP + j) * n + i] = row[j ^ sw];
        } else {
            mu[(int64_t)p * n + i] = NAN;
            al[(int64_t)p * n + i] = NAN;
            for (int j = 0; j < KP; ++j) vh[((int64_t)p * KP + j) * n + i] = NAN;
        }
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

static int gather_fits(pfb_engine* h, int cnt, const int32_t* d_units, double* mu, double* alpha, double* vh,
                       double* T, double* Vc, double* logdet, int32_t* jeff);
static int batch_download_impl(pfb_engine* h, pfb_elbo_out* o) {
    cudaStream_t st = h->stream;
    const size_t n = h->n, P = h->P, K = h->K, U = (size_t)h->U, KP = h->KP;
    PFB_D2H(o->elbo, h->dElbo.p, U * 8);
    PFB_D2H(o->elbo_se, h->dSe.p, U * 8);
    PFB_D2H(o->logp, h->dLogp.p, U * K * 8);
    PFB_D2H(o->logq, h->dLogq.p, U * K * 8);
    PFB_D2H(o->best_iter, h->dBestIter.p, P * 8);
    PFB_D2H(o->success, h->dSucc.p, P * 4);
    PFB_D2H(o->n_rejected, h->dRej.p, P * 8);
    if (o->draws) {
        int rcp = ensure_pool(h);
        if (rcp) return rcp;
    }
    PFB_D2H(o->draws, h->dPool.p, n * K * P * 8);
    PFB_D2H(o->draws_logp, h->dPoolLogp.p, K * P * 8);
    PFB_D2H(o->draws_logq, h->dPoolLogq.p, K * P * 8);
    if (o->all_draws) {
        if (!h->cfg.materialize_all) PFB_FAIL(h, PFB_ERR_STATE, "all_draws needs materialize_all = 1");
        PFB_D2H(o->all_draws, h->dAllDraws.p, n * K * U * 8);
    }
    const bool want_fit = o->fit_mu || o->fit_alpha || o->fit_vh || o->fit_T || o->fit_Vc || o->fit_logdet ||
                          o->fit_jeff;
    if (want_fit && P > 0)
        return gather_fits(h, (int)P, h->dBestUnit.as<int32_t>(), o->fit_mu, o->fit_alpha, o->fit_vh, o->fit_T,
                           o->fit_Vc, o->fit_logdet, o->fit_jeff);
    PFB_CUDA(h, cudaStreamSynchronize(st));
    return PFB_OK;
}

extern "C" int pfb_batch_download(pfb_handle h, pfb_elbo_out* o) {
    if (!h || !o) return PFB_ERR_ARG;
    if (!h->ran) PFB_FAIL(h, PFB_ERR_STATE, "pfb_batch_run has not been called");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    // optimistic: the copies are enqueued behind the kernels without waiting for the success flags;
    // only when a path did fail (known after the download's own synchronisation) is the pool part
    // fetched again after the fallback draws replaced it
    const bool was_pending = h->fail_pending;
    int rc = batch_download_impl(h, o);
    if (was_pending) {
        int r2 = resolve_failed(h);
        if (r2) return r2;
        if (h->n_failed > 0 && !h->have_normals && (o->draws || o->draws_logp || o->draws_logq))
            rc = batch_download_impl(h, o);
    }
    return rc;
}

// The fitted normals of `cnt` units (device list; < 0: NaN) in the reference's form, to host buffers
// (any may be NULL); synchronises the engine stream.
static int gather_fits(pfb_engine* h, int cnt, const int32_t* d_units, double* mu, double* alpha, double* vh,
                       double* T, double* Vc, double* logdet, int32_t* jeff) {
    cudaStream_t st = h->stream;
    const size_t n = h->n, KP = h->KP, P = (size_t)cnt;
    PFB_CUDA(h, h->dFitMu.ensure(n * P * 8));
    PFB_CUDA(h, h->dFitAlpha.ensure(n * P * 8));
    PFB_CUDA(h, h->dFitVh.ensure(n * KP * P * 8));
    PFB_CUDA(h, h->dFitT.ensure(KP * KP * P * 8));
    PFB_CUDA(h, h->dFitVc.ensure(KP * KP * P * 8));
    PFB_CUDA(h, h->dFitLogdet.ensure(P * 8));
    PFB_CUDA(h, h->dFitJeff.ensure(P * 4));
    if (h->generic) {
        PFB_CUDA(h, pfb_launch_kg_gather_fit(st, (int)P, (int)n, (int)KP, d_units, h->dFR.as<double>(), h->dHDR.as<double>(),
                                             h->dAlpha.as<double>(), h->dHistCnt.as<int32_t>(), h->dFitMu.as<double>(),
                                             h->dFitAlpha.as<double>(), h->dFitVh.as<double>(), h->dFitT.as<double>(),
                                             h->dFitVc.as<double>(), h->dFitLogdet.as<double>(),
                                             h->dFitJeff.as<int32_t>()));
    } else {
        pfb_gather_fit<<<(unsigned)P, 256, 0, st>>>((int)n, (int)KP, d_units, h->dFR2.as<double>(), h->dHDR.as<double>(),
                                                    h->dAlpha.as<double>(), h->dHistCnt.as<int32_t>(),
                                                    h->dFitMu.as<double>(), h->dFitAlpha.as<double>(),
                                                    h->dFitVh.as<double>(), h->dFitT.as<double>(),
                                                    h->dFitVc.as<double>(), h->dFitLogdet.as<double>(),
                                                    h->dFitJeff.as<int32_t>());
        PFB_CUDA(h, cudaGetLastError());
    }
    PFB_D2H(mu, h->dFitMu.p, n * P * 8);
    PFB_D2H(alpha, h->dFitAlpha.p, n * P * 8);
    PFB_D2H(vh, h->dFitVh.p, n * KP * P * 8);
    PFB_D2H(T, h->dFitT.p, KP * KP * P * 8);
    PFB_D2H(Vc, h->dFitVc.p, KP * KP * P * 8);
    PFB_D2H(logdet, h->dFitLogdet.p, P * 8);
    PFB_D2H(jeff, h->dFitJeff.p, P * 4);
    PFB_CUDA(h, cudaStreamSynchronize(st));
    return PFB_OK;
}

// Paths [p0, p1) of the device pool (the best-iteration draws left by pfb_batch_run, or the fresh
// draws of pfb_draw_from_fits(keep_as_pool)): lets a caller leave per-path draws on the device and
// fetch them only when they are looked at.
extern "C" int pfb_pool_download(pfb_handle h, int p0, int p1, double* draws, double* logp, double* logq) {
    if (!h) return PFB_ERR_ARG;
    if (!h->ran || h->poolK <= 0) PFB_FAIL(h, PFB_ERR_STATE, "no device pool (pfb_batch_run / pfb_draw_from_fits)");
    if (p0 < 0 || p1 < p0 || p1 > h->P) PFB_FAIL(h, PFB_ERR_ARG, "path range out of bounds");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    { int rf = resolve_failed(h); if (rf) return rf; }
    cudaStream_t st = h->stream;
    const size_t n = h->n, K = (size_t)h->poolK, cnt = (size_t)(p1 - p0);
    if (draws && cnt) {
        int rcp = ensure_pool(h);
        if (rcp) return rcp;
    }
    if (draws && cnt)
        PFB_CUDA(h, cudaMemcpyAsync(draws, h->dPool.as<double>() + (size_t)p0 * n * K, cnt * n * K * 8,
                                    cudaMemcpyDeviceToHost, st));
    if (logp && cnt)
        PFB_CUDA(h, cudaMemcpyAsync(logp, h->dPoolLogp.as<double>() + (size_t)p0 * K, cnt * K * 8,
                                    cudaMemcpyDeviceToHost, st));
    if (logq && cnt)
        PFB_CUDA(h, cudaMemcpyAsync(logq, h->dPoolLogq.as<double>() + (size_t)p0 * K, cnt * K * 8,
                                    cudaMemcpyDeviceToHost, st));
    PFB_CUDA(h, cudaStreamSynchronize(st));
    return PFB_OK;
}

extern "C" int pfb_elbo_batch(pfb_handle h, int n, int P, const int64_t* offsets, const double* positions,
                              const double* gradients, const uint64_t* seeds, const double* normals,
                              pfb_elbo_out* out) {
    // One call owns upload, compute and download, so the upload can be pipelined: the trajectories go
    // up in path groups on a copy stream and K1 / K2 of a group start when its points are resident
    // (the host buffers are not touched after the final synchronisation in pfb_batch_download).
    if (!h) return PFB_ERR_ARG;
    // parity mode / small batches (under 4 MB of trajectories the group-wise launches and events cost more
    // than the overlap returns: config 2 is 0.65 MB): the plain sequence
    if (normals || P < 8 || !offsets || n < 1 || (int64_t)offsets[P] * n * 16 < ((int64_t)4 << 20)) {
        int rc0 = pfb_batch_upload(h, n, P, offsets, positions, gradients, seeds, normals);
        if (rc0) return rc0;
        rc0 = pfb_batch_run(h);
        if (rc0) return rc0;
        return pfb_batch_download(h, out);
    }
    {
        const int64_t T = offsets[P], U = T - P;
        if (T > 0 && (!positions || !gradients)) PFB_FAIL(h, PFB_ERR_ARG, "positions / gradients are NULL");
        if (U > 0 && !seeds) PFB_FAIL(h, PFB_ERR_ARG, "seeds is NULL");
    }
    int rc = batch_prepare(h, n, P, offsets);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    if (!h->copy_stream) PFB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->k1_stream) PFB_CUDA(h, cudaStreamCreateWithFlags(&h->k1_stream, cudaStreamNonBlocking));
    for (auto& ev : h->up_ev)
        if (!ev) PFB_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : h->k1_ev)
        if (!ev) PFB_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (h->U > 0) PFB_CUDA(h, cudaMemcpyAsync(h->dSeeds.p, seeds, (size_t)h->U * 8, cudaMemcpyHostToDevice, st));
    h->have_normals = false;
    // groups of paths with about the same number of points
    const int NG = pfb_engine::kUpGroups;
    h->up_p[0] = 0;
    for (int g = 1; g < NG; ++g) {
        const int64_t want = offsets[P] * g / NG;
        int p = h->up_p[g - 1];
        while (p < P && offsets[p] < want) ++p;
        h->up_p[g] = p;
    }
    h->up_p[NG] = P;
    // the copy stream must not overwrite X / G while an earlier batch's kernels still read them
    PFB_CUDA(h, cudaEventRecord(h->up_ev[0], st));
    PFB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->up_ev[0], 0));
    PFB_CUDA(h, cudaStreamWaitEvent(h->k1_stream, h->up_ev[0], 0));
    for (int g = 0; g < NG; ++g) {
        const int64_t c0 = offsets[h->up_p[g]], c1 = offsets[h->up_p[g + 1]];
        const size_t bytes = (size_t)(c1 - c0) * (size_t)n * 8;
        if (bytes) {
            PFB_CUDA(h, cudaMemcpyAsync(h->dX.as<double>() + (size_t)c0 * n, positions + (size_t)c0 * n, bytes,
                                        cudaMemcpyHostToDevice, h->copy_stream));
            PFB_CUDA(h, cudaMemcpyAsync(h->dG.as<double>() + (size_t)c0 * n, gradients + (size_t)c0 * n, bytes,
                                        cudaMemcpyHostToDevice, h->copy_stream));
        }
        PFB_CUDA(h, cudaEventRecord(h->up_ev[g], h->copy_stream));
    }
    h->up_ngroups = NG;
    h->have_batch = true;
    rc = pfb_batch_run(h);
    if (rc) {
        h->up_ngroups = 0;
        cudaStreamSynchronize(h->copy_stream);
        return rc;
    }
    return pfb_batch_download(h, out);
}

extern "C" int pfb_pool_materialize(pfb_handle h) {
    if (!h) return PFB_ERR_ARG;
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    { int rf = resolve_failed(h); if (rf) return rf; }
    return ensure_pool(h);
}

// Multi-GPU resample: inds are 1-based indices into the GLOBAL pool (all ranks, run order); this
// engine owns [base, base + P K).  Its columns are written to d_out [n x m] (device); the others are
// left as they are (zero them first and sum-reduce across ranks).
extern "C" int pfb_pool_columns_device(pfb_handle h, int m, const void* d_inds, int64_t base, void* d_out) {
    if (!h || !d_inds || !d_out) return PFB_ERR_ARG;
    if (!h->ran || h->poolK <= 0) PFB_FAIL(h, PFB_ERR_STATE, "no device pool (pfb_batch_run / pfb_draw_from_fits)");
    if (h->poolK != h->K) PFB_FAIL(h, PFB_ERR_STATE, "column regeneration needs the pool of pfb_batch_run");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    { int rf = resolve_failed(h); if (rf) return rf; }
    return regen_columns(h, m, (const int64_t*)d_inds, base, (double*)d_out);
}

extern "C" int pfb_batch_device_view(pfb_handle h, pfb_device_view* v) {
    if (!h || !v) return PFB_ERR_ARG;
    if (!h->have_batch) PFB_FAIL(h, PFB_ERR_STATE, "no batch");
    { int rf = resolve_failed(h); if (rf) return rf; }  // the caller is about to read the pool's log densities
    // pool_draws is valid after pfb_pool_materialize (or a draws download); the log densities always are
    v->pool_draws = h->dPool.p;
    v->pool_logp = h->dPoolLogp.p;
    v->pool_logq = h->dPoolLogq.p;
    v->elbo = h->dElbo.p;
    v->stream = (void*)h->stream;
    v->n = h->n; v->K = h->K; v->P = h->P; v->U = h->U;
    return PFB_OK;
}

// K6 (PSIS) + the index draw K7 / K7b, enqueued on the engine stream.  d_pool != NULL: the selected
// columns are gathered from it into dOutDraws by the same kernels.
static int psis_enqueue(pfb_engine* h, int n, int64_t N, int K_run, const double* d_logp, const double* d_logq,
                        const double* d_logr, const double* d_pool, uint64_t seed, int ndraws, int importance,
                        int replace) {
    if (N < 1 || N > 2147483647LL) PFB_FAIL(h, PFB_ERR_SHAPE, "pool size out of range");
    if (K_run < 1 || ndraws < 0) PFB_FAIL(h, PFB_ERR_ARG, "bad K_run / ndraws");
    if (!replace && ndraws > N) PFB_FAIL(h, PFB_ERR_ARG, "Cannot draw more samples without replacement.");
    cudaStream_t st = h->stream;
    PFB_CUDA(h, h->dScal.ensure(pfb_psis_scalars_size()));
    PFB_CUDA(h, h->dInds.ensure((size_t)ndraws * 8 + 8));
    PFB_CUDA(h, h->dIds.ensure((size_t)ndraws * 8 + 8));
    PFB_CUDA(h, h->dOutDraws.ensure((size_t)n * ndraws * 8 + 8));
    if (importance) {
        // tail_length(r_eff = 1, S) = min(cld(S, 5), ceil(3 sqrt(S)));  grid m = 30 + floor(sqrt(M))
        const int M = (int)std::min<int64_t>((N + 4) / 5, (int64_t)ceil(3.0 * sqrt((double)N)));
        const int m_grid = 30 + (int)floor(sqrt((double)M));
        if (M + 1 > 8192) PFB_FAIL(h, PFB_ERR_UNSUPPORTED, "pool too large for the single-CTA PSIS kernel");
        PFB_CUDA(h, h->dLogw.ensure((size_t)N * 8));
        PFB_CUDA(h, h->dW.ensure((size_t)N * 8));
        PFB_CUDA(h, h->dCum.ensure((size_t)N * 8));
        const size_t wb = pfb_k6_workspace_bytes((int)N);
        PFB_CUDA(h, h->dSortWork.ensure(wb));
        PFB_CUDA(h, pfb_launch_k6(st, (int)N, M, m_grid, d_logp, d_logq, d_logr, h->dLogw.as<double>(),
                                  h->dW.as<double>(), h->dCum.as<uint64_t>(), h->dScal.p, h->dSortWork.p, wb));
    }
    if (replace) {
        PFB_CUDA(h, pfb_launch_k7(st, n, (int)N, K_run, seed, ndraws, importance ? h->dCum.as<uint64_t>() : nullptr,
                                  h->dScal.p, d_pool, h->dInds.as<int64_t>(), h->dIds.as<int64_t>(),
                                  d_pool ? h->dOutDraws.as<double>() : nullptr));
    } else {
        // K7b: exponential-key order statistics (weighted sampling without replacement)
        size_t tmp_bytes = 0;
        PFB_CUDA(h, pfb_k7b_temp_bytes((int)N, &tmp_bytes));
        PFB_CUDA(h, h->dSortWork.ensure((size_t)N * 24 + 64));
        PFB_CUDA(h, h->dSortTmp.ensure(tmp_bytes + 16));
        uint64_t* k_in = h->dSortWork.as<uint64_t>();
        uint64_t* k_out = k_in + N;
        int32_t* i_in = reinterpret_cast<int32_t*>(k_out + N);
        int32_t* i_out = i_in + N;
        PFB_CUDA(h, pfb_launch_k7b(st, n, (int)N, K_run, seed, ndraws, importance ? h->dLogw.as<double>() : nullptr,
                                   d_pool, k_in, k_out, i_in, i_out, h->dSortTmp.p, tmp_bytes, h->dInds.as<int64_t>(),
                                   h->dIds.as<int64_t>(), d_pool ? h->dOutDraws.as<double>() : nullptr));
    }
    return PFB_OK;
}

// Results of the last psis_enqueue to the caller's host buffers; one stream synchronisation.
static int psis_finish(pfb_engine* h, int n, int64_t N, int ndraws, int importance, pfb_resample_out* o,
                       bool have_draws) {
    cudaStream_t st = h->stream;
    psis_scalars_host sc;
    memset(&sc, 0, sizeof(sc));
    if (importance) {
        PFB_D2H(o->log_weights, h->dLogw.p, (size_t)N * 8);
        PFB_D2H(o->weights, h->dW.p, (size_t)N * 8);
        PFB_CUDA(h, cudaMemcpyAsync(&sc, h->dScal.p, sizeof(sc), cudaMemcpyDeviceToHost, st));
    }
    PFB_D2H(o->inds, h->dInds.p, (size_t)ndraws * 8);
    PFB_D2H(o->ids, h->dIds.p, (size_t)ndraws * 8);
    if (have_draws) PFB_D2H(o->draws, h->dOutDraws.p, (size_t)n * ndraws * 8);
    PFB_CUDA(h, cudaStreamSynchronize(st));
    if (o->pareto_k) *o->pareto_k = importance ? sc.pareto_k : NAN;
    if (o->tail_len) *o->tail_len = importance ? sc.tail_len : 0;
    if (importance && sc.Z == 0ull)
        PFB_FAIL(h, PFB_ERR_NUMERIC, "every importance weight is zero or undefined (all log ratios are NaN / -Inf): "
                                     "nothing to resample from");
    return PFB_OK;
}

static int psis_resample_impl(pfb_engine* h, int n, int64_t N, int K_run, const double* d_logp,
                              const double* d_logq, const double* d_logr, const double* d_pool, uint64_t seed,
                              int ndraws, int importance, int replace, pfb_resample_out* o, bool regen = false) {
    // regen: the engine's own pool, whose draws are not materialised: the selected columns are
    // regenerated by K3 (bit-identical to the pool's) instead of gathered
    const bool want_regen = regen && (o->draws != nullptr) && ndraws > 0;
    const bool want_draws = (o->draws != nullptr) && (d_pool != nullptr) && !want_regen;
    int rc = psis_enqueue(h, n, N, K_run, d_logp, d_logq, d_logr, want_draws ? d_pool : nullptr, seed, ndraws,
                          importance, replace);
    if (rc) return rc;
    if (want_regen) {
        int rcr = regen_columns(h, ndraws, h->dInds.as<int64_t>(), 0, h->dOutDraws.as<double>());
        if (rcr) return rcr;
    }
    return psis_finish(h, n, N, ndraws, importance, o, want_draws || want_regen);
}

extern "C" int pfb_psis_resample(pfb_handle h, uint64_t seed, int ndraws, int importance, int replace,
                                 pfb_resample_out* o) {
    if (!h || !o) return PFB_ERR_ARG;
    if (!h->ran || h->poolK <= 0) PFB_FAIL(h, PFB_ERR_STATE, "no device pool (pfb_batch_run / pfb_draw_from_fits)");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    // optimistic like pfb_batch_download: the whole PSIS chain is enqueued behind the ELBO stage (no host
    // round trip between pfb_batch_run and here); a failed path, known after the chain's own
    // synchronisation, repeats it on the corrected pool
    const bool was_pending = h->fail_pending;
    int rc = psis_resample_impl(h, h->n, (int64_t)h->poolP * h->poolK, h->poolK, h->dPoolLogp.as<double>(),
                                h->dPoolLogq.as<double>(), nullptr, h->dPool.as<double>(), seed, ndraws,
                                importance, replace, o, /*regen=*/!h->pool_ready);
    if (was_pending) {
        int r2 = resolve_failed(h);
        if (r2) return r2;
        if (h->n_failed > 0 && !h->have_normals)
            rc = psis_resample_impl(h, h->n, (int64_t)h->poolP * h->poolK, h->poolK, h->dPoolLogp.as<double>(),
                                    h->dPoolLogq.as<double>(), nullptr, h->dPool.as<double>(), seed, ndraws,
                                    importance, replace, o, /*regen=*/!h->pool_ready);
    }
    return rc;
}

extern "C" int pfb_psis_resample_device(pfb_handle h, int n, int64_t N, int K_run, const void* d_logp,
                                        const void* d_logq, const void* d_pool, uint64_t seed, int ndraws,
                                        int importance, int replace, pfb_resample_out* o) {
    if (!h || !o) return PFB_ERR_ARG;
    if (importance && (!d_logp || !d_logq)) PFB_FAIL(h, PFB_ERR_ARG, "d_logp / d_logq are NULL");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    return psis_resample_impl(h, n, N, K_run, (const double*)d_logp, (const double*)d_logq, nullptr,
                              (const double*)d_pool, seed, ndraws, importance, replace, o);
}

extern "C" int pfb_psis_resample_host(pfb_handle h, int n, int64_t N, int K_run, const double* log_ratios,
                                      const double* pool, uint64_t seed, int ndraws, int importance, int replace,
                                      pfb_resample_out* o) {
    if (!h || !o) return PFB_ERR_ARG;
    if (N < 1) PFB_FAIL(h, PFB_ERR_SHAPE, "empty pool");
    if (importance && !log_ratios) PFB_FAIL(h, PFB_ERR_ARG, "log_ratios is NULL");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const double* d_logr = nullptr;
    const double* d_pool = nullptr;
    if (importance) {
        PFB_CUDA(h, h->dTmpLogr.ensure((size_t)N * 8));
        PFB_CUDA(h, cudaMemcpyAsync(h->dTmpLogr.p, log_ratios, (size_t)N * 8, cudaMemcpyHostToDevice, st));
        d_logr = h->dTmpLogr.as<double>();
    }
    if (pool && o->draws) {
        PFB_CUDA(h, h->dTmpPool.ensure((size_t)n * N * 8));
        PFB_CUDA(h, cudaMemcpyAsync(h->dTmpPool.p, pool, (size_t)n * N * 8, cudaMemcpyHostToDevice, st));
        d_pool = h->dTmpPool.as<double>();
    }
    return psis_resample_impl(h, n, N, K_run, nullptr, nullptr, d_logr, d_pool, seed, ndraws, importance, replace, o);
}

// ---- multi-GPU: the PSIS pool exchange behind the ABI (src/multipath.jl:190-225) --------------
// Paths shard over the GPUs with no communication until the pool.  The exchange is the lean form of
// the "all-gather of the pool": every rank all-gathers the per-draw log densities (16 B per pool draw),
// runs PSIS and the index draw replicated (deterministic kernels + counter RNG => identical on every
// rank), regenerates the selected columns it owns, and the ranks sum-reduce the n x ndraws result.
// NCCL is loaded lazily (dlopen), so the library has no link-time dependency on it: inside a PyTorch
// process this resolves to the NCCL torch already loaded, elsewhere to the system libnccl.so.2.
#include <dlfcn.h>

namespace {
struct pfb_nccl_uid { char internal[128]; };
typedef void* pfb_nccl_comm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(pfb_nccl_uid*) = nullptr;
    int (*CommInitRank)(pfb_nccl_comm_t*, int, pfb_nccl_uid, int) = nullptr;
    int (*CommDestroy)(pfb_nccl_comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, pfb_nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, pfb_nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, pfb_nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

NcclApi* nccl_api() {
    static NcclApi api;
    if (api.lib) return &api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.err = std::string("cannot load NCCL: ") + dlerror();
        return nullptr;
    }
#define PFB_NCCL_SYM(field, name)                                             \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name)); \
    if (!api.field) {                                                         \
        api.err = std::string("NCCL symbol missing: ") + name;               \
        api.lib = nullptr;                                                    \
        return nullptr;                                                       \
    }
    PFB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    PFB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    PFB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PFB_NCCL_SYM(AllGather, "ncclAllGather")
    PFB_NCCL_SYM(AllReduce, "ncclAllReduce")
    PFB_NCCL_SYM(Broadcast, "ncclBroadcast")
    PFB_NCCL_SYM(GroupStart, "ncclGroupStart")
    PFB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PFB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PFB_NCCL_SYM
    return &api;
}
}  // namespace

#define PFB_NCCL(h, expr)                                                                          \
    do {                                                                                           \
        int r_ = (expr);                                                                           \
        if (r_ != 0) {                                                                             \
            (h)->err = std::string(#expr) + ": NCCL error " + nccl_api()->GetErrorString(r_);      \
            return 1000 + r_;                                                                      \
        }                                                                                          \
    } while (0)

extern "C" int pfb_comm_unique_id(void* id128) {
    if (!id128) return PFB_ERR_ARG;
    NcclApi* a = nccl_api();
    if (!a) {
        g_create_err = "NCCL is not available (libnccl.so.2 could not be loaded)";
        return PFB_ERR_UNSUPPORTED;
    }
    pfb_nccl_uid u;
    int r = a->GetUniqueId(&u);
    if (r != 0) return 1000 + r;
    memcpy(id128, &u, sizeof(u));
    return PFB_OK;
}

extern "C" int pfb_comm_destroy(pfb_handle h) {
    if (!h) return PFB_ERR_ARG;
    if (h->comm) {
        cudaSetDevice(h->cfg.device);
        cudaStreamSynchronize(h->stream);
        nccl_api()->CommDestroy(h->comm);
    }
    h->comm = nullptr;
    h->comm_world = 1;
    h->comm_rank = 0;
    return PFB_OK;
}

// One handle per GPU; every rank calls with the same id (from pfb_comm_unique_id on one of them).
extern "C" int pfb_comm_init(pfb_handle h, const void* id128, int rank, int world) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return PFB_ERR_ARG;
    NcclApi* a = nccl_api();
    if (!a) PFB_FAIL(h, PFB_ERR_UNSUPPORTED, "NCCL is not available (libnccl.so.2 could not be loaded)");
    pfb_comm_destroy(h);
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    pfb_nccl_uid u;
    memcpy(&u, id128, sizeof(u));
    pfb_nccl_comm_t c = nullptr;
    PFB_NCCL(h, a->CommInitRank(&c, world, u, rank));
    h->comm = c;
    h->comm_world = world;
    h->comm_rank = rank;
    return PFB_OK;
}

// Single process, several GPUs (what a Julia caller without MPI does): handle i is rank i.
extern "C" int pfb_comm_init_all(pfb_handle* hs, int nh) {
    if (!hs || nh < 1) return PFB_ERR_ARG;
    for (int i = 0; i < nh; ++i)
        if (!hs[i]) return PFB_ERR_ARG;
    NcclApi* a = nccl_api();
    if (!a) PFB_FAIL(hs[0], PFB_ERR_UNSUPPORTED, "NCCL is not available (libnccl.so.2 could not be loaded)");
    pfb_nccl_uid u;
    PFB_NCCL(hs[0], a->GetUniqueId(&u));
    for (int i = 0; i < nh; ++i) pfb_comm_destroy(hs[i]);
    std::vector<pfb_nccl_comm_t> cs((size_t)nh, nullptr);
    PFB_NCCL(hs[0], a->GroupStart());
    for (int i = 0; i < nh; ++i) {
        PFB_CUDA(hs[i], cudaSetDevice(hs[i]->cfg.device));
        PFB_NCCL(hs[i], a->CommInitRank(&cs[(size_t)i], nh, u, i));
    }
    PFB_NCCL(hs[0], a->GroupEnd());
    for (int i = 0; i < nh; ++i) {
        hs[i]->comm = cs[(size_t)i];
        hs[i]->comm_world = nh;
        hs[i]->comm_rank = i;
    }
    return PFB_OK;
}

// out[t] = column inds[t] of this engine's materialised pool if it owns it (global 1-based index in
// [base + 1, base + cnt]), else untouched
__global__ void pfb_gather_owned_columns(int n, int m, const int64_t* __restrict__ inds, int64_t base, int64_t cnt,
                                         const double* __restrict__ pool, double* __restrict__ out) {
    const int t = blockIdx.x;
    const int64_t j = inds[t] - 1 - base;
    if (j < 0 || j >= cnt) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[(int64_t)t * n + i] = pool[j * n + i];
}

// The exchange for `nh` handles driven by this thread (nh = 1: one process per GPU; nh = world: one
// process, all GPUs).  Every collective is issued for all local handles inside one NCCL group.
static int pool_exchange_impl(pfb_engine** hs, int nh, const int32_t* paths_per_rank, uint64_t seed, int ndraws,
                              int importance, int replace, pfb_resample_out** outs) {
    NcclApi* a = nccl_api();
    pfb_engine* h0 = hs[0];
    if (!a) PFB_FAIL(h0, PFB_ERR_UNSUPPORTED, "NCCL is not available");
    const int world = h0->comm_world;
    int K_run = 0;
    for (int i = 0; i < nh; ++i) {
        pfb_engine* h = hs[i];
        if (!h->comm || h->comm_world != world) PFB_FAIL(h, PFB_ERR_STATE, "pfb_comm_init has not been called");
        if (!h->ran) PFB_FAIL(h, PFB_ERR_STATE, "no device pool (pfb_batch_run / pfb_draw_from_fits)");
        { int rf = resolve_failed(h); if (rf) return rf; }
        if (paths_per_rank[h->comm_rank] != h->poolP) PFB_FAIL(h, PFB_ERR_SHAPE, "paths_per_rank[rank] differs from this engine's pool");
        if (h->poolP > 0 && h->poolK <= 0) PFB_FAIL(h, PFB_ERR_STATE, "no device pool (pfb_batch_run / pfb_draw_from_fits)");
        if (h->poolK > 0) {  // (a rank without runs states its draws per run through pfb_pool_set(P = 0, K_run))
            if (K_run && K_run != h->poolK) PFB_FAIL(h, PFB_ERR_SHAPE, "draws per run differ between engines");
            K_run = h->poolK;
        }
    }
    if (K_run == 0) K_run = h0->K;  // (a rank without runs: every rank uses ndraws_elbo draws per run)
    std::vector<int64_t> off((size_t)world + 1, 0);
    bool equal = true;
    for (int r = 0; r < world; ++r) {
        if (paths_per_rank[r] < 0) PFB_FAIL(h0, PFB_ERR_ARG, "negative paths_per_rank");
        off[(size_t)r + 1] = off[(size_t)r] + (int64_t)paths_per_rank[r] * K_run;
        equal = equal && paths_per_rank[r] == paths_per_rank[0];
    }
    const int64_t N = off[(size_t)world];
    if (N < 1) PFB_FAIL(h0, PFB_ERR_SHAPE, "empty pool");
    const int n = h0->n;
    for (int i = 0; i < nh; ++i) {
        pfb_engine* h = hs[i];
        PFB_CUDA(h, cudaSetDevice(h->cfg.device));
        PFB_CUDA(h, h->dGLogp.ensure((size_t)N * 8));
        PFB_CUDA(h, h->dGLogq.ensure((size_t)N * 8));
        PFB_CUDA(h, h->dPoolLogp.ensure(8));  // (ranks without runs still need valid send pointers)
        PFB_CUDA(h, h->dPoolLogq.ensure(8));
    }
    if (importance) {
        // C1: all-gather of the per-draw log densities (ragged shards: one broadcast per owner)
        PFB_NCCL(h0, a->GroupStart());
        for (int i = 0; i < nh; ++i) {
            pfb_engine* h = hs[i];
            PFB_CUDA(h, cudaSetDevice(h->cfg.device));
            if (equal) {
                const size_t cnt = (size_t)(off[1] - off[0]);
                PFB_NCCL(h, a->AllGather(h->dPoolLogp.p, h->dGLogp.p, cnt, kNcclFloat64, h->comm, h->stream));
                PFB_NCCL(h, a->AllGather(h->dPoolLogq.p, h->dGLogq.p, cnt, kNcclFloat64, h->comm, h->stream));
            } else {
                for (int r = 0; r < world; ++r) {
                    const size_t cnt = (size_t)(off[(size_t)r + 1] - off[(size_t)r]);
                    if (cnt == 0) continue;
                    double* gp = h->dGLogp.as<double>() + off[(size_t)r];
                    double* gq = h->dGLogq.as<double>() + off[(size_t)r];
                    const bool mine = (r == h->comm_rank);
                    PFB_NCCL(h, a->Broadcast(mine ? h->dPoolLogp.p : (const void*)gp, gp, cnt, kNcclFloat64, r, h->comm, h->stream));
                    PFB_NCCL(h, a->Broadcast(mine ? h->dPoolLogq.p : (const void*)gq, gq, cnt, kNcclFloat64, r, h->comm, h->stream));
                }
            }
        }
        PFB_NCCL(h0, a->GroupEnd());
    }
    // K6 + K7 replicated; every rank writes the selected columns it owns into a zeroed n x ndraws buffer
    for (int i = 0; i < nh; ++i) {
        pfb_engine* h = hs[i];
        PFB_CUDA(h, cudaSetDevice(h->cfg.device));
        int rc = psis_enqueue(h, n, N, K_run, importance ? h->dGLogp.as<double>() : nullptr,
                              importance ? h->dGLogq.as<double>() : nullptr, nullptr, nullptr, seed, ndraws, importance,
                              replace);
        if (rc) return rc;
        if (ndraws > 0) {
            PFB_CUDA(h, cudaMemsetAsync(h->dOutDraws.p, 0, (size_t)n * ndraws * 8, h->stream));
            const int64_t base = off[(size_t)h->comm_rank];
            if (h->poolP > 0 && h->pool_ready) {
                pfb_gather_owned_columns<<<ndraws, 128, 0, h->stream>>>(n, ndraws, h->dInds.as<int64_t>(), base,
                                                                        (int64_t)h->poolP * K_run, h->dPool.as<double>(),
                                                                        h->dOutDraws.as<double>());
                PFB_CUDA(h, cudaGetLastError());
            } else if (h->poolP > 0) {
                if (h->poolK != h->K || h->poolP != h->P)
                    PFB_FAIL(h, PFB_ERR_STATE, "column regeneration needs the pool of pfb_batch_run");
                rc = regen_columns(h, ndraws, h->dInds.as<int64_t>(), base, h->dOutDraws.as<double>());
                if (rc) return rc;
            }
        }
    }
    if (ndraws > 0) {
        // C2: sum-reduce of the n x ndraws result (every column is non-zero on exactly one rank)
        PFB_NCCL(h0, a->GroupStart());
        for (int i = 0; i < nh; ++i) {
            pfb_engine* h = hs[i];
            PFB_CUDA(h, cudaSetDevice(h->cfg.device));
            PFB_NCCL(h, a->AllReduce(h->dOutDraws.p, h->dOutDraws.p, (size_t)n * ndraws, kNcclFloat64, kNcclSum, h->comm,
                                     h->stream));
        }
        PFB_NCCL(h0, a->GroupEnd());
    }
    for (int i = 0; i < nh; ++i) {
        pfb_engine* h = hs[i];
        PFB_CUDA(h, cudaSetDevice(h->cfg.device));
        int rc = psis_finish(h, n, N, ndraws, importance, outs[i], outs[i]->draws != nullptr && ndraws > 0);
        if (rc) return rc;
    }
    return PFB_OK;
}

// Replaces _compute_psis_result + _resample (src/multipath.jl:220-225) over the runs of ALL ranks:
// paths_per_rank[world] = runs owned by each rank (rank order = run order, so the pool keeps the
// reference's component order, src/multipath.jl:217); every rank passes the same seed / ndraws and
// receives the same indices, ids, weights and draws.
extern "C" int pfb_pool_exchange_resample(pfb_handle h, const int32_t* paths_per_rank, uint64_t seed, int ndraws,
                                          int importance, int replace, pfb_resample_out* out) {
    if (!h || !paths_per_rank || !out) return PFB_ERR_ARG;
    pfb_engine* hs[1] = {h};
    pfb_resample_out* os[1] = {out};
    return pool_exchange_impl(hs, 1, paths_per_rank, seed, ndraws, importance, replace, os);
}
extern "C" int pfb_pool_exchange_resample_all(pfb_handle* hs, int nh, const int32_t* paths_per_rank, uint64_t seed,
                                              int ndraws, int importance, int replace, pfb_resample_out* outs) {
    if (!hs || nh < 1 || !paths_per_rank || !outs) return PFB_ERR_ARG;
    std::vector<pfb_resample_out*> os((size_t)nh);
    for (int i = 0; i < nh; ++i) {
        if (!hs[i]) return PFB_ERR_ARG;
        os[(size_t)i] = &outs[i];
    }
    return pool_exchange_impl(hs, nh, paths_per_rank, seed, ndraws, importance, replace, os.data());
}

// A device pool from HOST arrays (P runs, K_run draws each): draws[n x K_run x P], logp / logq
// [K_run x P].  For pools assembled on the host (top-up draws beyond ndraws_elbo, retried paths, a rank
// without runs: P = 0) that then take part in pfb_psis_resample / pfb_pool_exchange_resample.
extern "C" int pfb_pool_set(pfb_handle h, int P, int K_run, const double* draws, const double* logp, const double* logq) {
    if (!h || P < 0 || K_run < 1 || (P > 0 && (!logp || !logq))) return PFB_ERR_ARG;
    if (h->model < 0) PFB_FAIL(h, PFB_ERR_STATE, "no model registered");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    { int rf = resolve_failed(h); if (rf) return rf; }
    cudaStream_t st = h->stream;
    const size_t n = h->model_n, Ps = (size_t)P, K = (size_t)K_run;
    if (!draws && P > 0 && (K_run != h->K || P != h->P || !h->ran))
        PFB_FAIL(h, PFB_ERR_ARG, "a pool without draws must be the current batch's (its draws are regenerated)");
    PFB_CUDA(h, h->dPoolLogp.ensure(K * Ps * 8 + 8));
    PFB_CUDA(h, h->dPoolLogq.ensure(K * Ps * 8 + 8));
    if (P > 0) {
        PFB_CUDA(h, cudaMemcpyAsync(h->dPoolLogp.p, logp, K * Ps * 8, cudaMemcpyHostToDevice, st));
        PFB_CUDA(h, cudaMemcpyAsync(h->dPoolLogq.p, logq, K * Ps * 8, cudaMemcpyHostToDevice, st));
    }
    if (draws) {
        PFB_CUDA(h, h->dPool.ensure(n * K * Ps * 8 + 8));
        if (P > 0) PFB_CUDA(h, cudaMemcpyAsync(h->dPool.p, draws, n * K * Ps * 8, cudaMemcpyHostToDevice, st));
    }
    PFB_CUDA(h, cudaStreamSynchronize(st));
    h->n = (int)n;
    h->poolK = K_run;
    h->poolP = P;
    h->pool_ready = draws != nullptr;
    h->ran = true;
    return PFB_OK;
}

// Page-lock / unlock a caller-owned host buffer (cudaHostRegister): output buffers that a caller
// reuses across batches then receive their device-to-host copies at full PCIe rate.
extern "C" int pfb_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return PFB_ERR_ARG;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return PFB_OK;
    }
    return e == cudaSuccess ? PFB_OK : (int)e;
}
extern "C" int pfb_host_unregister(void* p) {
    if (!p) return PFB_ERR_ARG;
    cudaError_t e = cudaHostUnregister(p);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return PFB_OK;
    }
    return e == cudaSuccess ? PFB_OK : (int)e;
}

extern "C" int pfb_get_timings(pfb_handle h, double* ms6) {
    if (!h || !ms6) return PFB_ERR_ARG;
    if (!h->ran) PFB_FAIL(h, PFB_ERR_STATE, "no batch has run");
    PFB_CUDA(h, cudaSetDevice(h->cfg.device));
    PFB_CUDA(h, cudaEventSynchronize(h->ev[5]));
    float t;
    for (int i = 0; i < 5; ++i) {
        PFB_CUDA(h, cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]));
        ms6[i] = t;
    }
    PFB_CUDA(h, cudaEventElapsedTime(&t, h->ev[0], h->ev[5]));
    ms6[5] = t;
    return h->launches;
}
