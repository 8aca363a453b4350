/* A caller of libpfb200.so written in plain C99 — no Python, no torch, nothing but include/pfb200.h.
 * It does what the Julia shim (julia/PathfinderB200.jl) does through ccall:
 *   pfb_create -> pfb_register_model -> pfb_elbo_batch -> pfb_psis_resample -> pfb_destroy
 * which replaces fit_mvnormals / maximize_elbo (src/singlepath.jl:301-308) and
 * _compute_psis_result / _resample (src/multipath.jl:215-225) of the reference.
 *
 * Usage: abi_client [--inputs-only] <dump-file>
 * Builds P = 3 synthetic optimiser traces in n = 24 dimensions (iso-normal target), runs the
 * path and dumps inputs and outputs (native endianness) so that tests/test_c_client.py can push
 * the SAME inputs through the ctypes binding and compare bit for bit.
 * Exit codes: 0 OK, 2 usage / IO, 3 the engine could not be created (no CUDA device: the product
 * path has no CPU fallback), 4 any later ABI error. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pfb200.h"

#define N 24
#define P 3
#define K 64
#define J 6
#define NDRAWS 40

/* 53-bit uniform in [0, 1) from a 64-bit LCG: deterministic inputs without any library */
static double lcg(uint64_t* s) {
    *s = *s * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(*s >> 11) * (1.0 / 9007199254740992.0);
}

static int fail(pfb_handle h, const char* what, int rc) {
    fprintf(stderr, "abi_client: %s failed with code %d: %s\n", what, rc, h ? pfb_last_error(h) : "(no handle)");
    return 4;
}

int main(int argc, char** argv) {
    static const int64_t L[P] = {5, 3, 7};
    int64_t offsets[P + 1];
    int64_t T, U, t;
    int p, i, rc;
    double *X, *G, *elbo, *se, *draws, *w;
    uint64_t* seeds;
    int64_t best_iter[P], inds[NDRAWS], ids[NDRAWS], tail_len = 0;
    int32_t success[P];
    double pareto_k = 0.0;
    pfb_config cfg;
    pfb_handle h = NULL;
    pfb_elbo_out out;
    pfb_resample_out ro;
    FILE* f;
    int32_t hdr[8];

    const int inputs_only = argc == 3 && strcmp(argv[1], "--inputs-only") == 0;
    const char* path = inputs_only ? argv[2] : argv[1];
    if (argc != 2 && !inputs_only) {
        fprintf(stderr, "usage: abi_client [--inputs-only] <dump-file>\n");
        return 2;
    }
    offsets[0] = 0;
    for (p = 0; p < P; ++p) offsets[p + 1] = offsets[p] + L[p] + 1;
    T = offsets[P];
    U = T - P;
    X = (double*)malloc(sizeof(double) * N * T);
    G = (double*)malloc(sizeof(double) * N * T);
    seeds = (uint64_t*)malloc(sizeof(uint64_t) * U);
    elbo = (double*)malloc(sizeof(double) * U);
    se = (double*)malloc(sizeof(double) * U);
    draws = (double*)malloc(sizeof(double) * N * NDRAWS);
    w = (double*)malloc(sizeof(double) * K * P);
    if (!X || !G || !seeds || !elbo || !se || !draws || !w) return 2;

    /* a smooth descent on a quadratic with the SPD Hessian H = diag(d) + B B' (B: N x 3), as an optimiser
     * would record it: points x_l and gradients of the LOG density, -H x_l.  (The trajectory need not
     * belong to the registered target: the engine fits the normals to whatever trace it is given.) */
    {
        uint64_t st = 0x243F6A8885A308D3ull;
        double d[N], B[N][3], hx[N];
        int k;
        for (i = 0; i < N; ++i) {
            d[i] = 0.5 + 1.5 * lcg(&st);
            for (k = 0; k < 3; ++k) B[i][k] = lcg(&st) - 0.5;
        }
        for (p = 0; p < P; ++p) {
            double x[N];
            for (i = 0; i < N; ++i) x[i] = 2.0 * lcg(&st) - 1.0;
            for (t = offsets[p]; t < offsets[p + 1]; ++t) {
                double dot[3] = {0.0, 0.0, 0.0};
                const double step = (0.05 + 0.25 * lcg(&st)) / 8.0;
                for (k = 0; k < 3; ++k)
                    for (i = 0; i < N; ++i) dot[k] += B[i][k] * x[i];
                for (i = 0; i < N; ++i) {
                    hx[i] = d[i] * x[i] + B[i][0] * dot[0] + B[i][1] * dot[1] + B[i][2] * dot[2];
                    X[(size_t)t * N + i] = x[i];
                    G[(size_t)t * N + i] = -hx[i];
                }
                for (i = 0; i < N; ++i) x[i] = x[i] - step * hx[i] + 0.01 * (lcg(&st) - 0.5);
            }
        }
    }
    for (t = 0; t < U; ++t) seeds[t] = 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1) + 12345u;

    if (inputs_only) { /* the synthetic inputs alone (CPU tests run the oracle on them) */
        f = fopen(path, "wb");
        if (!f) return 2;
        hdr[0] = N; hdr[1] = P; hdr[2] = K; hdr[3] = J; hdr[4] = NDRAWS; hdr[5] = (int32_t)T; hdr[6] = (int32_t)U; hdr[7] = 1;
        fwrite(hdr, sizeof(hdr), 1, f);
        fwrite(offsets, sizeof(int64_t), P + 1, f);
        fwrite(X, sizeof(double), (size_t)N * T, f);
        fwrite(G, sizeof(double), (size_t)N * T, f);
        fwrite(seeds, sizeof(uint64_t), (size_t)U, f);
        fclose(f);
        return 0;
    }
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = 0;
    cfg.history_length = J;
    cfg.ndraws_elbo = K;
    cfg.eps = 1e-12;
    rc = pfb_create(&h, &cfg);
    if (rc != PFB_OK) {
        fprintf(stderr, "abi_client: pfb_create failed with code %d (no CUDA device? the engine has no CPU fallback)\n", rc);
        return 3;
    }
    rc = pfb_register_model(h, PFB_MODEL_ISONORMAL, N, NULL, 0);
    if (rc) return fail(h, "pfb_register_model", rc);

    memset(&out, 0, sizeof(out));
    out.elbo = elbo;
    out.elbo_se = se;
    out.best_iter = best_iter;
    out.success = success;
    rc = pfb_elbo_batch(h, N, P, offsets, X, G, seeds, NULL, &out);
    if (rc) return fail(h, "pfb_elbo_batch", rc);

    memset(&ro, 0, sizeof(ro));
    ro.weights = w;
    ro.pareto_k = &pareto_k;
    ro.tail_len = &tail_len;
    ro.inds = inds;
    ro.ids = ids;
    ro.draws = draws;
    rc = pfb_psis_resample(h, 2024u, NDRAWS, 1, 1, &ro);
    if (rc) return fail(h, "pfb_psis_resample", rc);
    rc = pfb_destroy(h);
    if (rc) return fail(NULL, "pfb_destroy", rc);

    f = fopen(path, "wb");
    if (!f) return 2;
    hdr[0] = N; hdr[1] = P; hdr[2] = K; hdr[3] = J; hdr[4] = NDRAWS; hdr[5] = (int32_t)T; hdr[6] = (int32_t)U; hdr[7] = 0;
    fwrite(hdr, sizeof(hdr), 1, f);
    fwrite(offsets, sizeof(int64_t), P + 1, f);
    fwrite(X, sizeof(double), (size_t)N * T, f);
    fwrite(G, sizeof(double), (size_t)N * T, f);
    fwrite(seeds, sizeof(uint64_t), (size_t)U, f);
    fwrite(elbo, sizeof(double), (size_t)U, f);
    fwrite(se, sizeof(double), (size_t)U, f);
    fwrite(best_iter, sizeof(int64_t), P, f);
    fwrite(success, sizeof(int32_t), P, f);
    fwrite(w, sizeof(double), (size_t)K * P, f);
    fwrite(&pareto_k, sizeof(double), 1, f);
    fwrite(&tail_len, sizeof(int64_t), 1, f);
    fwrite(inds, sizeof(int64_t), NDRAWS, f);
    fwrite(ids, sizeof(int64_t), NDRAWS, f);
    fwrite(draws, sizeof(double), (size_t)N * NDRAWS, f);
    fclose(f);
    printf("abi_client: %d paths, %d units, best ELBO of path 0 = %.17g, pareto k = %.6f\n", P, (int)U,
           elbo[best_iter[0] > 0 ? best_iter[0] - 1 : 0], pareto_k);
    free(X); free(G); free(seeds); free(elbo); free(se); free(draws); free(w);
    return 0;
}
