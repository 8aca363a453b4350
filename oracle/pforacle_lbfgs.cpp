// oracle/pforacle_lbfgs.cpp — TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
//
// CPU side of the engine's L-BFGS trajectory contract (pathfinder_b200/csrc/pf_lbfgs.h, SURVEY §8
// row f1; the trajectory producer of src/optimize.jl:35-59).  The algorithm source is the shared
// header; this file supplies the host execution context, which EMULATES the device's reduction
// order (PF_LBFGS_T strided partial sums, xor-butterfly inside each group of 32, then a butterfly
// over the group totals), so the trajectories are bit-identical to kernel K0's.
// Built by oracle/Makefile with -ffp-contract=off.
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <vector>

#include "../pathfinder_b200/csrc/pf_lbfgs.h"

namespace {

inline double butterfly32(double* v) {
    for (int off = 16; off > 0; off >>= 1) {
        double nv[32];
        for (int l = 0; l < 32; ++l) nv[l] = v[l] + v[l ^ off];
        for (int l = 0; l < 32; ++l) v[l] = nv[l];
    }
    return v[0];
}

inline double block_tree(const double* part) {
    const int NW = PF_LBFGS_T / 32;
    double tot[32];
    for (int l = 0; l < 32; ++l) tot[l] = 0.0;
    for (int w = 0; w < NW; ++w) {
        double v[32];
        for (int l = 0; l < 32; ++l) v[l] = part[w * 32 + l];
        tot[w] = butterfly32(v);
    }
    return butterfly32(tot);
}

struct HostCtx {
    int n;
    template <class F>
    void each(F f) {
        for (int i = 0; i < n; ++i) f(i);
    }
    template <class F>
    void each_n(int count, F f) {
        for (int i = 0; i < count; ++i) f(i);
    }
    // one sequential fma chain over the columns per row; column-outer so the matrix streams once
    void matvec_cols(int nr, int nc, const double* A, const double* v, double init, double* out) {
        for (int i = 0; i < nr; ++i) out[i] = init;
        for (int j = 0; j < nc; ++j) {
            const double* col = A + (size_t)j * nr;
            const double vj = v[j];
            for (int i = 0; i < nr; ++i) out[i] = fma(col[i], vj, out[i]);
        }
    }
    template <class F>
    double sum(F f) {
        return sum_n(n, f);
    }
    template <class F>
    double sum_n(int count, F f) {
        double part[PF_LBFGS_T];
        for (int t = 0; t < PF_LBFGS_T; ++t) {
            double acc = 0.0;
            for (int i = t; i < count; i += PF_LBFGS_T) acc = f(i, acc);
            part[t] = acc;
        }
        return block_tree(part);
    }
    template <class F>
    void sum2(F f, double& a, double& b) {
        double pa[PF_LBFGS_T], pb[PF_LBFGS_T];
        for (int t = 0; t < PF_LBFGS_T; ++t) {
            double x = 0.0, y = 0.0;
            for (int i = t; i < n; i += PF_LBFGS_T) f(i, x, y);
            pa[t] = x;
            pb[t] = y;
        }
        a = block_tree(pa);
        b = block_tree(pb);
    }
    template <class F>
    double maxv(F f) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc = fmax(acc, f(i));
        return acc;
    }
    void sync() {}
};

}  // namespace

// One path.  X, G: n x max_points column-major, FX[max_points].  mp0 / mp1: DIAGNORMAL mean and
// 1 / sd; DENSENORMAL mean and precision (n x n column-major); HLOGISTIC X (nobs x (n-2)) and y, mp2 = X' ((n-2) x nobs).  Returns the number of recorded points.
extern "C" int pfo_lbfgs_path(int family, int n, int nobs, const double* mp0, const double* mp1, const double* mp2,
                              double mc0, int J,
                              int maxiters, int max_points, double gtol, double ftol, const double* x0, double* X,
                              double* G, double* FX, int* status, int* nevals) {
    std::vector<double> ws((size_t)(2 * J + 1) * n + (size_t)(nobs > n ? nobs : n));
    pf_lbfgs_model m{family, n, mp0, mp1, mc0, ws.data() + (size_t)(2 * J + 1) * n, nobs, mp2};
    pf_lbfgs_opts o{J, maxiters, max_points, gtol, ftol};
    HostCtx c{n};
    return pf_lbfgs_run(c, m, o, x0, X, G, FX, ws.data(), status, nevals);
}
