// pf_lbfgs.h — the engine's L-BFGS trajectory contract, shared source for the
// device kernel (k0_lbfgs.cu, one CTA per path) and the CPU oracle (oracle/pforacle_lbfgs.cpp).
//
// Replaces, for the registered closed-form target families, the trajectory producer
//   optimize_with_trace(prob, optimizer; maxiters, fail_on_nonfinite)          src/optimize.jl:35-59
//   default_optimizer = Optim.LBFGS(m = history_length, HagerZhang line search) src/Pathfinder.jl:29-35
// and records what OptimizationCallback records per iteration (x, log density, gradient of the log
// density; src/optimize.jl:94-101), starting with the initial point, and stops — after recording
// the point — on a non-finite value (src/optimize.jl:103-105).
//
// Optim.jl / LineSearches.jl are third-party Julia code absent from /root/reference (parity
// unpinned), so the iterates of THIS contract are not Optim's: it is the textbook
// algorithm — two-loop recursion with H0 = (s'y / y'y) I (Nocedal & Wright Alg. 7.4/7.5) and a
// strong-Wolfe bracketing / zoom line search with safeguarded cubic interpolation (Alg. 3.5/3.6,
// c1 = 1e-4, c2 = 0.9).  What IS pinned is CPU == GPU: every reduction goes through the execution
// context `Ctx`, whose device and host implementations add in the same order (PF_LBFGS_T strided
// partial sums, then a fixed butterfly), and exp() is pf_math.h's, so both sides produce
// bit-identical trajectories (tests/test_gpu_parity.py::test_device_lbfgs_*).
//
// Ctx interface (i runs over the elements owned by the caller: all of them on the host, the
// thread's stride on the device):
//   template <class F> void   each(F f)     f(i)
//   template <class F> double sum(F f)      acc = f(i, acc), acc starts at 0; block-wide total
//   template <class F> void   sum2(F f, double& a, double& b)   f(i, a, b) updates both partials
//   template <class F> double maxv(F f)     max of f(i) >= 0 (0 when empty)
//   each_n(count, f) / sum_n(count, f)      the same over an index range other than the dimension
//   matvec_cols(nr, nc, A, v, init, out)    out[i] = init + sum_j A[i + j nr] v[j], ONE sequential fma
//                                           chain over j per row (column-major A): which thread owns a
//                                           row, and how many rows it interleaves, does not change a bit
//   void sync()                             make element writes visible to every caller (a context whose
//                                           reductions and element loops have different owners also
//                                           synchronises at the start of every reduction)
// Build contract: -fmad=false / -ffp-contract=off (pf_math.h).
#pragma once
#include <float.h>
#include "pf_math.h"

#define PF_LBFGS_T 256      // threads per path = strided partial sums of the reduction contract
#define PF_LBFGS_MAXJ 12    // history_length limit of the engine (pfb_kp_of)
#define PF_LBFGS_MAXLS 40   // function evaluations per line search

// family ids of include/pfb200.h
#define PF_LBFGS_ISONORMAL 0
#define PF_LBFGS_FUNNEL 1
#define PF_LBFGS_DIAGNORMAL 2
#define PF_LBFGS_DENSENORMAL 3
#define PF_LBFGS_HLOGISTIC 4

#define PF_LBFGS_CONVERGED_G 0   // max |grad| <= gtol
#define PF_LBFGS_CONVERGED_F 1   // relative decrease <= ftol
#define PF_LBFGS_MAXITER 2       // maxiters or the point capacity reached
#define PF_LBFGS_LINESEARCH 3    // no acceptable step
#define PF_LBFGS_NONFINITE 4     // non-finite log density / gradient (src/optimize.jl:103-105)

struct pf_lbfgs_model {
    int family;
    int n;
    const double* p0;  // DIAGNORMAL / DENSENORMAL: mean[n]
    const double* p1;  // DIAGNORMAL: 1 / sd[n];  DENSENORMAL: precision P[n x n], column-major (symmetric)
    double c0;         // DIAGNORMAL: -sum(log sd) - n/2 log(2 pi);  HLOGISTIC: the prior's constant
    double* zbuf;      // DENSENORMAL: n doubles of per-path scratch (x - mean);  HLOGISTIC: nobs doubles
    int nobs;          // HLOGISTIC: observations; p0 = X[nobs x (n-2)] column-major, p1 = y[nobs]
    const double* p2;  // HLOGISTIC: X' [(n-2) x nobs] column-major (the same numbers, transposed)
};

struct pf_lbfgs_opts {
    int J;           // history_length
    int maxiters;    // src/optimize.jl:40
    int max_points;  // capacity of the trajectory slab (columns), >= 1
    double gtol;     // Optim g_abstol
    double ftol;     // relative objective decrease
};

// log density, its gradient (into glog) and max |gradient| (INFINITY if any entry is not finite)
template <class Ctx>
PF_HD void pf_lbfgs_eval(Ctx& c, const pf_lbfgs_model& m, const double* x, double* glog, double& logp,
                         double& gmax) {
    const int n = m.n;
    c.sync();  // x[0] may have been written by another caller
    if (m.family == PF_LBFGS_FUNNEL) {
        // docs/src/examples/quickstart.md:229-234
        const double x0 = x[0];
        const double ss = c.sum([&](int i, double a) { return i > 0 ? fma(x[i], x[i], a) : a; });
        const double e = pf_exp(-x0);
        const double t3 = x0 / 3.0;
        logp = (fma(t3, t3, (double)(n - 1) * x0) + e * ss) / -2.0;
        const double g0 = (((2.0 * x0) / 9.0 + (double)(n - 1)) - e * ss) / -2.0;
        c.each([&](int i) { glog[i] = (i == 0) ? g0 : -(e * x[i]); });
    } else if (m.family == PF_LBFGS_DENSENORMAL) {
        // docs/src/examples/quickstart.md:17-24: log p = -(x - m)' P (x - m) / 2, gradient -P (x - m).
        // Row i of the product is one sequential fma chain over the columns (the order the oracle
        // repeats); consecutive rows sit in consecutive threads, so every column step is coalesced.
        double* z = m.zbuf;
        c.each([&](int i) { z[i] = x[i] - m.p0[i]; });
        c.sync();
        c.matvec_cols(n, n, m.p1, z, 0.0, glog);  // glog = P z for now
        c.sync();
        const double q = c.sum([&](int i, double a) { return fma(z[i], glog[i], a); });
        logp = q / -2.0;
        c.each([&](int i) { glog[i] = -glog[i]; });
    } else if (m.family == PF_LBFGS_HLOGISTIC) {
        // theta = (log tau, b0, b_1..b_p); log tau ~ N(0,1), b0 ~ N(0, 2.5^2),
        // b_j ~ N(0, tau^2), y_i ~ Bernoulli(sigmoid(b0 + x_i'b)).  eta: one sequential fma chain per
        // observation (coalesced column sweep); X'r: one sequential chain per coefficient.
        const int p = n - 2, nobs = m.nobs;
        const double lt = x[0], b0 = x[1];
        const double* Xm = m.p0;
        const double* yv = m.p1;
        double* r = m.zbuf;
        c.matvec_cols(nobs, p, Xm, x + 2, b0, r);  // eta = b0 + X b
        c.sync();
        // log-likelihood sum_i y eta - log(1 + exp(eta)), evaluated without overflow
        const double ll = c.sum_n(nobs, [&](int i, double a) {
            const double eta = r[i];
            const double l1pe = (eta > 0.0) ? eta + pf_log1p(pf_exp(-eta)) : pf_log1p(pf_exp(eta));
            return a + (yv[i] * eta - l1pe);
        });
        c.sync();
        c.each_n(nobs, [&](int i) {
            const double eta = r[i];
            double sg;
            if (eta >= 0.0) {
                sg = 1.0 / (1.0 + pf_exp(-eta));
            } else {
                const double e = pf_exp(eta);
                sg = e / (1.0 + e);
            }
            r[i] = yv[i] - sg;
        });
        c.sync();
        const double rsum = c.sum_n(nobs, [&](int i, double a) { return a + r[i]; });
        const double bb = c.sum([&](int i, double a) { return i >= 2 ? fma(x[i], x[i], a) : a; });
        const double e2 = pf_exp(-2.0 * lt);
        logp = ((((-0.5 * lt) * lt - 0.5 * ((b0 / 2.5) * (b0 / 2.5))) - (0.5 * bb) * e2) - (double)p * lt) + m.c0 + ll;
        c.sync();
        c.matvec_cols(p, nobs, m.p2, r, 0.0, glog + 2);  // X'r: one chain over the observations per coefficient
        c.sync();
        c.each([&](int i) {
            if (i == 0) {
                glog[i] = (bb * e2 - lt) - (double)p;
            } else if (i == 1) {
                glog[i] = rsum - b0 / 6.25;
            } else {
                glog[i] = glog[i] - x[i] * e2;
            }
        });
    } else if (m.family == PF_LBFGS_DIAGNORMAL) {
        const double ss = c.sum([&](int i, double a) {
            const double z = (x[i] - m.p0[i]) * m.p1[i];
            return fma(z, z, a);
        });
        logp = fma(ss, -0.5, m.c0);
        c.each([&](int i) { glog[i] = -(((x[i] - m.p0[i]) * m.p1[i]) * m.p1[i]); });
    } else {
        const double ss = c.sum([&](int i, double a) { return fma(x[i], x[i], a); });
        logp = ss / -2.0;
        c.each([&](int i) { glog[i] = -x[i]; });
    }
    gmax = c.maxv([&](int i) {
        const double a = fabs(glog[i]);
        return (a <= DBL_MAX) ? a : (double)INFINITY;
    });
}

// Runs one path.  X, G: n x max_points column-major slabs (points, gradients of the LOG density),
// FX[max_points] log densities; ws: (2 J + 1) n doubles (m.zbuf: n more for the dense normal).  Returns the number of points recorded
// (L + 1 >= 1); *status = PF_LBFGS_*, *nevals = density evaluations.
template <class Ctx>
PF_HD int pf_lbfgs_run(Ctx& c, const pf_lbfgs_model& m, const pf_lbfgs_opts& o, const double* x0, double* X,
                       double* G, double* FX, double* ws, int* status, int* nevals) {
    const int n = m.n, J = o.J;
    double* d = ws;
    double* S = ws + n;
    double* Y = S + (size_t)J * n;
    double rho[PF_LBFGS_MAXJ], aj[PF_LBFGS_MAXJ];
    const double C1 = 1e-4, C2 = 0.9;

    c.each([&](int i) { X[i] = x0[i]; });
    double logp, gmax;
    pf_lbfgs_eval(c, m, X, G, logp, gmax);
    FX[0] = logp;
    int npts = 1, nev = 1, st = PF_LBFGS_MAXITER;
    double f = -logp;
    if (!(fabs(f) <= DBL_MAX) || gmax == (double)INFINITY) {
        *status = PF_LBFGS_NONFINITE;
        *nevals = nev;
        return npts;
    }
    if (gmax <= o.gtol) {
        *status = PF_LBFGS_CONVERGED_G;
        *nevals = nev;
        return npts;
    }
    int hist = 0, head = 0;
    double gamma = 1.0;
    for (int k = 0; k < o.maxiters && npts < o.max_points; ++k) {
        const double* xk = X + (size_t)k * n;
        const double* gk = G + (size_t)k * n;  // gradient of log p; the minimised f = -log p has -gk
        double* xt = X + (size_t)(k + 1) * n;
        double* gt = G + (size_t)(k + 1) * n;
        // ---- direction d = -H (-gk): two-loop recursion --------------------------------------
        c.each([&](int i) { d[i] = -gk[i]; });
        for (int jj = 0; jj < hist; ++jj) {
            const int j = (head - 1 - jj + 2 * J) % J;
            const double* Sj = S + (size_t)j * n;
            const double* Yj = Y + (size_t)j * n;
            const double a = rho[j] * c.sum([&](int i, double acc) { return fma(Sj[i], d[i], acc); });
            aj[jj] = a;
            c.each([&](int i) { d[i] = fma(-a, Yj[i], d[i]); });
        }
        if (hist > 0) c.each([&](int i) { d[i] = gamma * d[i]; });
        for (int jj = hist - 1; jj >= 0; --jj) {
            const int j = (head - 1 - jj + 2 * J) % J;
            const double* Sj = S + (size_t)j * n;
            const double* Yj = Y + (size_t)j * n;
            const double b = rho[j] * c.sum([&](int i, double acc) { return fma(Yj[i], d[i], acc); });
            const double w = aj[jj] - b;
            c.each([&](int i) { d[i] = fma(w, Sj[i], d[i]); });
        }
        c.each([&](int i) { d[i] = -d[i]; });
        double dg = c.sum([&](int i, double acc) { return fma(-gk[i], d[i], acc); });
        if (!(dg < 0.0)) {  // not a descent direction: restart from steepest descent
            hist = 0;
            c.each([&](int i) { d[i] = gk[i]; });
            dg = -c.sum([&](int i, double acc) { return fma(gk[i], gk[i], acc); });
            if (!(dg < 0.0)) {
                st = PF_LBFGS_LINESEARCH;
                break;
            }
        }
        double a_init = 1.0;
        if (hist == 0) {
            const double gn = sqrt(c.sum([&](int i, double acc) { return fma(gk[i], gk[i], acc); }));
            a_init = (gn > 1.0) ? 1.0 / gn : 1.0;
        }
        // ---- strong-Wolfe line search on phi(a) = f(xk + a d) ----------------------------------
        double phi = 0.0, dphi = 0.0, gmx = 0.0, lp_t = 0.0, a_last = -1.0;
        auto evalat = [&](double a) {
            c.each([&](int i) { xt[i] = fma(a, d[i], xk[i]); });
            pf_lbfgs_eval(c, m, xt, gt, lp_t, gmx);
            ++nev;
            dphi = c.sum([&](int i, double acc) { return fma(-gt[i], d[i], acc); });
            phi = -lp_t;
            if (!(fabs(phi) <= DBL_MAX)) phi = (double)INFINITY;  // NaN / Inf: never acceptable
            a_last = a;
        };
        const double phi0 = f, dphi0 = dg;
        double a_lo = 0.0, phi_lo = phi0, dphi_lo = dphi0;
        double a_hi = 0.0, phi_hi = phi0, dphi_hi = dphi0;
        double a_acc = -1.0;
        bool bracket = false;
        int ls = 0;
        {
            double a = a_init, a_prev = 0.0, phi_prev = phi0, dphi_prev = dphi0;
            while (ls < PF_LBFGS_MAXLS) {
                evalat(a);
                ++ls;
                if (phi > fma(C1 * a, dphi0, phi0) || (ls > 1 && phi >= phi_prev)) {
                    a_lo = a_prev; phi_lo = phi_prev; dphi_lo = dphi_prev;
                    a_hi = a; phi_hi = phi; dphi_hi = dphi;
                    bracket = true;
                    break;
                }
                if (fabs(dphi) <= -C2 * dphi0) {
                    a_acc = a;
                    break;
                }
                if (dphi >= 0.0) {
                    a_lo = a; phi_lo = phi; dphi_lo = dphi;
                    a_hi = a_prev; phi_hi = phi_prev; dphi_hi = dphi_prev;
                    bracket = true;
                    break;
                }
                a_prev = a; phi_prev = phi; dphi_prev = dphi;
                a_lo = a; phi_lo = phi; dphi_lo = dphi;  // best Armijo point so far (fallback)
                a = 4.0 * a;
            }
        }
        while (bracket && a_acc < 0.0 && ls < PF_LBFGS_MAXLS) {
            const double w = a_hi - a_lo;
            // minimiser of the cubic through (a_lo, phi_lo, dphi_lo), (a_hi, phi_hi, dphi_hi)
            const double d1 = (dphi_lo + dphi_hi) - 3.0 * ((phi_lo - phi_hi) / (a_lo - a_hi));
            const double rad = fma(d1, d1, -(dphi_lo * dphi_hi));
            const double d2 = (w > 0.0) ? sqrt(rad) : -sqrt(rad);
            double at = a_hi - w * (((dphi_hi + d2) - d1) / ((dphi_hi - dphi_lo) + 2.0 * d2));
            const double lo_b = (a_lo < a_hi) ? a_lo : a_hi, hi_b = (a_lo < a_hi) ? a_hi : a_lo;
            const double margin = 0.1 * (hi_b - lo_b);
            if (!(at >= lo_b + margin && at <= hi_b - margin)) at = fma(0.5, w, a_lo);  // also NaN
            if (at == a_lo || at == a_hi) break;  // the bracket has collapsed
            evalat(at);
            ++ls;
            if (phi > fma(C1 * at, dphi0, phi0) || phi >= phi_lo) {
                a_hi = at; phi_hi = phi; dphi_hi = dphi;
            } else {
                if (fabs(dphi) <= -C2 * dphi0) {
                    a_acc = at;
                    break;
                }
                if (dphi * (a_hi - a_lo) >= 0.0) {
                    a_hi = a_lo; phi_hi = phi_lo; dphi_hi = dphi_lo;
                }
                a_lo = at; phi_lo = phi; dphi_lo = dphi;
            }
        }
        if (a_acc < 0.0) {
            // evaluation budget spent: take the best sufficient-decrease point, if there is one
            if (a_lo > 0.0 && phi_lo <= fma(C1 * a_lo, dphi0, phi0) && phi_lo < phi0) {
                a_acc = a_lo;
            } else {
                st = PF_LBFGS_LINESEARCH;
                break;
            }
        }
        if (a_acc != a_last) evalat(a_acc);  // make column k+1 hold the accepted point
        // ---- record the iteration (src/optimize.jl:94-101) -----------------------------------
        FX[k + 1] = lp_t;
        ++npts;
        const double f_new = phi;
        if (gmx == (double)INFINITY) {
            st = PF_LBFGS_NONFINITE;
            break;
        }
        // ---- history update ------------------------------------------------------------------
        {
            double* Sj = S + (size_t)head * n;
            double* Yj = Y + (size_t)head * n;
            c.each([&](int i) {
                Sj[i] = xt[i] - xk[i];
                Yj[i] = gk[i] - gt[i];
            });
            double sy = 0.0, yy = 0.0;
            c.sum2([&](int i, double& a, double& b) {
                a = fma(Sj[i], Yj[i], a);
                b = fma(Yj[i], Yj[i], b);
            }, sy, yy);
            if (sy > 0.0 && yy > 0.0 && sy <= DBL_MAX && yy <= DBL_MAX) {
                rho[head] = 1.0 / sy;
                gamma = sy / yy;
                head = (head + 1) % J;
                if (hist < J) ++hist;
            }
        }
        if (gmx <= o.gtol) {
            st = PF_LBFGS_CONVERGED_G;
            break;
        }
        {
            double sc = fabs(f) > fabs(f_new) ? fabs(f) : fabs(f_new);
            if (sc < 1.0) sc = 1.0;
            if (f - f_new <= o.ftol * sc) {
                st = PF_LBFGS_CONVERGED_F;
                break;
            }
        }
        f = f_new;
    }
    *status = st;
    *nevals = nev;
    return npts;
}
