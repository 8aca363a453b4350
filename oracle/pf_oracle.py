"""CPU oracle for the ELBO-and-resample hot path of Pathfinder.jl  —  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module.  The product path (``pathfinder_b200``) never does.

This is a float64 NumPy/SciPy restatement of the reference algorithm (Pathfinder.jl v0.10.7);
every function cites the reference lines it follows (paths relative to the reference repo).
Linear algebra goes through the *same LAPACK routines* Julia dispatches to
(``dgeqrt``/``dgemqrt`` for ``qr`` and ``lmul!(Q, .)``, ``dpotrf`` for ``cholesky``,
``dtrtrs`` for triangular solves), so Householder sign conventions match the reference's.

Parity status:
  * pinned by the reference's own fixtures / known answers: ``lbfgs_inverse_hessian`` (literal
    S0/Y0 fixture, test/inverse_hessian.jl:19-44), Woodbury identities vs dense
    (test/woodbury.jl), ``mu = theta + Sigma g`` (test/mvnormal.jl:28), analytic ELBO
    (test/elbo.jl:8-28), ``_findmax_skipnan`` table (test/utils.jl:6-13), log-ratio ordering and
    resample membership (test/resample.jl).
  * PARITY UNPINNED (third-party code absent from the reference tree, no golden vectors):
    the normal stream (Julia Random), PSIS smoothed weights / Pareto k (PSIS.jl 0.2-0.9),
    StatsBase's weighted-sampling index stream.  For these the oracle restates the published
    algorithm and the engine's own RNG contract (pathfinder_b200/csrc/pf_rng.h).
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field

import numpy as np
from scipy.linalg import lapack

_HERE = os.path.dirname(os.path.abspath(__file__))
LOG2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------------
# C helper (bit-reproducible math + RNG contract)
# --------------------------------------------------------------------------------------------
_clib = None


def clib():
    """Load oracle/_build/libpforacle.so (built by oracle/Makefile)."""
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "_build", "libpforacle.so")
        if not os.path.exists(path):
            import subprocess

            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = ctypes.CDLL(path)
        lib.pfo_normals.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.pfo_resample_bits.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p]
        lib.pfo_mulhi64.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        lib.pfo_mulhi64.restype = ctypes.c_uint64
        for name in ("pfo_exp_v", "pfo_log_v", "pfo_log1p_v", "pfo_expm1_v"):
            getattr(lib, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        _clib = lib
    return _clib


def _vec(name, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    getattr(clib(), name)(x.ctypes.data, y.ctypes.data, x.size)
    return y


def pf_exp(x):
    return _vec("pfo_exp_v", x)


def pf_log(x):
    return _vec("pfo_log_v", x)


def pf_log1p(x):
    return _vec("pfo_log1p_v", x)


def pf_expm1(x):
    return _vec("pfo_expm1_v", x)


def contract_normals(seed: int, n: int, K: int) -> np.ndarray:
    """u[n, K] (Fortran order, i.e. column k is draw k) of the engine's RNG contract."""
    u = np.empty((K, n), dtype=np.float64)
    clib().pfo_normals(ctypes.c_uint64(int(seed) & (2**64 - 1)), n, K, u.ctypes.data)
    return u.T  # (n, K), column-major storage


# --------------------------------------------------------------------------------------------
# L-BFGS inverse-Hessian reconstruction     (reference: src/inverse_hessian.jl)
# --------------------------------------------------------------------------------------------
def gilbert_init(alpha, s, y):
    """Diagonal initial inverse Hessian, Gilbert & Lemarechal eq 4.9.
    reference: src/inverse_hessian.jl:5-10"""
    a = np.dot(y, alpha * y)
    b = np.dot(y, s)
    c = np.dot(s, s / alpha)
    return b / (a / alpha + y**2 - (a / c) * (s / alpha) ** 2)


def lbfgs_inverse_hessian(alpha, S0, Y0, history_ind, history_length):
    """Byrd et al. (1994) compact representation H = diag(alpha) + B D B'.
    reference: src/inverse_hessian.jl:98-133.  history_ind is 1-based (0 = empty)."""
    J = history_length
    n = alpha.shape[0]
    B = np.zeros((n, 2 * J), order="F")
    D = np.zeros((2 * J, 2 * J), order="F")
    if J == 0:
        return B, D
    hist = list(range(history_ind, J)) + list(range(0, history_ind))  # :105, 0-based
    S = S0[:, hist]
    Y = Y0[:, hist]
    B[:, :J] = alpha[:, None] * Y  # :117
    B[:, J:] = S  # :118
    R = np.triu(S.T @ Y)  # :119-121
    nRinv, info = lapack.dtrtrs(R, -np.eye(J), lower=0)  # :122-124
    if info != 0:
        nRinv = np.full((J, J), np.nan)
    D[:J, J:] = nRinv
    D[J:, :J] = nRinv.T  # :125
    M = np.diag(np.diag(R)) + Y.T @ B[:, :J]  # :126-128  E + Y' H0 Y
    M = np.triu(M) + np.triu(M, 1).T  # copytri! 'U'
    D[J:, J:] = nRinv.T @ M @ nRinv  # :129-130
    return B, D


@dataclass
class WoodburyPD:
    """W = diag(alpha) + B D B' with its square-root factorisation W = R'R,
    R = diag(Vc, I) Q' U,  U = diag(sqrt(alpha)).     reference: src/woodbury.jl:201-207"""

    alpha: np.ndarray
    B: np.ndarray
    D: np.ndarray
    k: int = 0
    sqrt_alpha: np.ndarray = field(default=None, repr=False)
    Vh: np.ndarray = field(default=None, repr=False)  # n x k reflectors (unit lower trapezoid)
    T: np.ndarray = field(default=None, repr=False)  # k x k compact-WY factor
    Rq: np.ndarray = field(default=None, repr=False)  # k x k upper
    Vc: np.ndarray = field(default=None, repr=False)  # k x k upper Cholesky factor of I + Rq D Rq'
    qr_raw: np.ndarray = field(default=None, repr=False)
    pd_ok: bool = True

    @property
    def n(self):
        return self.alpha.shape[0]

    def logdet(self):
        """reference: src/woodbury.jl:77-80 (2 (logdet U + logdet V))"""
        ld = np.sum(np.log(self.alpha))
        if self.k:
            ld += 2.0 * np.sum(np.log(np.diag(self.Vc)))
        return ld

    def _q_apply(self, x, trans):
        if self.k == 0:
            return x
        c = np.asfortranarray(x if x.ndim == 2 else x[:, None])
        v = np.asfortranarray(self.qr_raw[:, : self.k])  # the k = min(n, ncols) reflectors
        out, info = lapack.dgemqrt(v, self.T, c, side="L", trans="T" if trans else "N")
        assert info == 0
        return out if x.ndim == 2 else out[:, 0]

    def lmul_L(self, x):
        """x <- L x,  L = U' Q diag(Vc', I).     reference: src/woodbury.jl:136-143"""
        x = np.array(x, dtype=np.float64, copy=True)
        k = self.k
        if k:
            x[:k] = self.Vc.T @ x[:k]
        x = self._q_apply(x, trans=False)
        return (self.sqrt_alpha * x.T).T

    def lmul_R(self, x):
        """x <- R x,  R = diag(Vc, I) Q' U.      reference: src/woodbury.jl:129-135"""
        x = (self.sqrt_alpha * np.array(x, dtype=np.float64).T).T
        x = self._q_apply(x, trans=True)
        k = self.k
        if k:
            x = np.array(x, copy=True)
            x[:k] = self.Vc @ x[:k]
        return x

    def ldiv_L(self, x):
        """x <- L \\ x.                            reference: src/woodbury.jl:158-165"""
        x = (np.array(x, dtype=np.float64).T / self.sqrt_alpha).T
        x = self._q_apply(x, trans=True)
        k = self.k
        if k:
            x = np.array(x, copy=True)
            sol, info = lapack.dtrtrs(self.Vc, x[:k], lower=0, trans=1)
            x[:k] = sol
        return x

    def mul(self, x):
        """W x = L (R x).                         reference: src/woodbury.jl:346-349, :64-68"""
        return self.lmul_L(self.lmul_R(x))

    def invquad(self, x):
        """column-wise x' W^-1 x.                 reference: src/woodbury.jl:378-382, :425-436"""
        v = self.ldiv_L(x)
        return np.sum(v * v, axis=0)

    def dense(self):
        return np.diag(self.alpha) + self.B @ self.D @ self.B.T


def pdfactorize(alpha, B, D) -> WoodburyPD:
    """reference: src/woodbury.jl:201-207 with A = Diagonal(alpha)."""
    n, k = B.shape
    # k below = min(size(U, 1), size(V, 1)) of src/woodbury.jl:130: the number of reflectors
    W = WoodburyPD(alpha=np.asarray(alpha, dtype=np.float64), B=B, D=D, k=min(n, k))
    W.sqrt_alpha = np.sqrt(W.alpha)  # U = cholesky(Diagonal).U
    if k == 0:
        return W
    A = np.asfortranarray(B / W.sqrt_alpha[:, None])  # U' \ B
    # Julia: qr(A) -> LAPACK.geqrt!(A, min(min(m, n), 36)); here k <= 36 so one block.
    nb = min(min(n, k), 36)
    qr_raw, T, info = lapack.dgeqrt(nb, A)
    assert info == 0
    W.qr_raw = qr_raw
    W.T = T
    kk = min(n, k)
    W.Rq = np.triu(qr_raw[:kk, :])
    Vh = np.tril(qr_raw[:, :kk], -1)
    Vh[np.arange(kk), np.arange(kk)] = 1.0
    W.Vh = Vh
    C = np.eye(kk) + W.Rq @ D @ W.Rq.T  # muladd(R, D * R', I)
    C = np.triu(C) + np.triu(C, 1).T  # Symmetric(.) reads the upper triangle
    Vc, info = lapack.dpotrf(C, lower=0)
    if info != 0:
        # reference throws PosDefException here (src/woodbury.jl:205); the engine turns this
        # into "this iteration's ELBO is NaN" (SURVEY §8b error convention).
        W.pd_ok = False
        Vc = np.full((kk, kk), np.nan)
    W.Vc = np.triu(Vc)
    return W


def nocedal_wright_scaling(alpha, s, y):
    """Hinit used by test/inverse_hessian.jl:50: fill(y's / y'y)."""
    return np.full_like(alpha, np.dot(y, s) / np.dot(y, y))


def lbfgs_inverse_hessians(thetas, grads, history_length=6, eps=1e-12, hinit=None):
    """reference: src/inverse_hessian.jl:25-66 (hinit = Hinit keyword, default gilbert_init).  thetas, grads: (n, L+1) arrays (columns = points;
    grads are gradients of the log density).  Returns (list of WoodburyPD length L+1,
    num_bfgs_updates_rejected, per-point state list)."""
    thetas = np.asarray(thetas, dtype=np.float64)
    grads = np.asarray(grads, dtype=np.float64)
    n, Lp1 = thetas.shape
    L = Lp1 - 1
    J = history_length
    history_ind = 0
    history_eff = 0
    S = np.zeros((n, min(J, L)), order="F") if L > 0 else np.zeros((n, 0))
    Y = np.zeros_like(S)
    alpha = np.ones(n)
    Hs = []
    states = []
    B, D = lbfgs_inverse_hessian(alpha, S, Y, history_ind, history_eff)
    Hs.append(pdfactorize(alpha.copy(), B, D))
    states.append((history_ind, history_eff))
    rejected = 0
    theta = thetas[:, 0]
    g = grads[:, 0]
    for l in range(1, L + 1):
        theta1, g1 = thetas[:, l], grads[:, l]
        s = theta1 - theta  # :45
        y = g - g1  # :46
        if np.dot(y, s) > eps * np.sum(y * y):  # :47
            history_ind = history_ind % J + 1  # mod1(ind + 1, J)
            history_eff = max(history_ind, history_eff)
            S[:, history_ind - 1] = s
            Y[:, history_ind - 1] = y
            alpha = (hinit or gilbert_init)(alpha, s, y)  # :55
        else:
            rejected += 1
        theta, g = theta1, g1
        B, D = lbfgs_inverse_hessian(alpha, S, Y, history_ind, history_eff)
        Hs.append(pdfactorize(alpha.copy(), B, D))
        states.append((history_ind, history_eff))
    return Hs, rejected, states


def fit_mvnormals(thetas, grads, history_length=6, eps=1e-12):
    """reference: src/mvnormal.jl:14-21.  Returns (mus (n, L+1), list of WoodburyPD, rejected)."""
    Hs, rejected, _ = lbfgs_inverse_hessians(thetas, grads, history_length, eps)
    mus = np.empty_like(np.asarray(thetas, dtype=np.float64))
    for l, W in enumerate(Hs):
        mus[:, l] = thetas[:, l] + W.mul(grads[:, l])  # muladd(Sigma, grad, theta)
    return mus, Hs, rejected


# --------------------------------------------------------------------------------------------
# sampling, ELBO                               (reference: src/mvnormal.jl, src/elbo.jl)
# --------------------------------------------------------------------------------------------
def rand_and_logpdf(u, mu, W: WoodburyPD):
    """reference: src/mvnormal.jl:24-39 with the normals u (n, K) supplied by the caller."""
    n = mu.shape[0]
    unormsq = np.sum(u * u, axis=0)  # :31
    x = W.lmul_L(u) + mu[:, None]  # :32-33
    logq = (n * LOG2PI + W.logdet() + unormsq) / -2.0  # :36
    return x, logq


def elbo_and_samples(u, logp_fn, mu, W):
    """reference: src/elbo.jl:12-20.  logp_fn maps (n, K) -> (K,)."""
    x, logq = rand_and_logpdf(u, mu, W)
    if not W.pd_ok:
        logq = np.full_like(logq, np.nan)
    logp = logp_fn(x)
    logr = logp - logq
    K = logr.shape[0]
    elbo = np.mean(logr)
    se = math.sqrt(np.sum((logr - elbo) ** 2) / (K - 1) / K) if K > 1 else float("nan")
    return dict(value=elbo, std_err=se, draws=x, logp=logp, logq=logq, logr=logr)


def findmax_skipnan(values):
    """reference: src/utils.jl:57-72.  Returns (max, 1-based index); (nan, 0) if empty."""
    state = None
    for i, x in enumerate(values, start=1):
        if state is None:
            state = (x, i)
            continue
        if math.isnan(x):
            continue
        if math.isnan(state[0]) or x > state[0]:
            state = (x, i)
    return state if state is not None else (float("nan"), 0)


def maximize_elbo(seeds, logp_fn, mus, Hs, K, normals=None):
    """reference: src/elbo.jl:1-10 applied to dists[2:end] (src/singlepath.jl:306-308):
    estimate l (1-based) uses (mus[:, l], Hs[l]) and seed seeds[l-1]."""
    L = len(Hs) - 1
    ests = []
    for l in range(1, L + 1):
        u = normals[l - 1] if normals is not None else contract_normals(seeds[l - 1], mus.shape[0], K)
        ests.append(elbo_and_samples(u, logp_fn, mus[:, l], Hs[l]))
    if not ests:
        return 0, ests
    _, lopt = findmax_skipnan([e["value"] for e in ests])
    return lopt, ests


def failed_path_draws(fallback_seed, mu, W, K):
    """reference: src/singlepath.jl:224-228.  A failed path returns rand(rng, fit_distributions[
    fit_iteration + 1], ndraws): fresh draws from the fit of the "best" iteration (the identity fit of
    iteration 0 when there is none) with the PATH's rng — here the engine's contract normals of the
    path's fallback seed.  Returns (draws [n, K], logq [K])."""
    u = contract_normals(int(fallback_seed), mu.shape[0], K)
    return rand_and_logpdf(np.asarray(u), mu, W)


def path_success(L, ests, lopt):
    """reference: src/singlepath.jl:297-314"""
    if L <= 0 or not ests:
        return False
    e = ests[lopt - 1]["value"]
    return (not math.isnan(e)) and e != -math.inf


# --------------------------------------------------------------------------------------------
# target log densities of the registered device-side model family
# --------------------------------------------------------------------------------------------
def logp_isonormal(x):
    """test/singlepath.jl:15   logp(x) = -sum(abs2, x) / 2"""
    return -0.5 * np.sum(x * x, axis=0)


def logp_funnel(x):
    """docs/src/examples/quickstart.md:229-234"""
    n = x.shape[0]
    tau = x[0]
    ss = np.sum(x[1:] ** 2, axis=0)
    return ((tau / 3.0) ** 2 + (n - 1) * tau + np.exp(-tau) * ss) / -2.0


def make_logp_dense_gaussian(mean, prec):
    """docs/src/examples/quickstart.md:17-24   -(x-m)' P (x-m) / 2"""

    def f(x):
        z = x - mean[:, None]
        return -0.5 * np.sum(z * (prec @ z), axis=0)

    return f


def make_logp_hier_logistic(X, y):
    """SURVEY §8d config 4 (not in the reference; spec fixed there): theta = (log tau, b0, b),
    log tau ~ N(0,1), b0 ~ N(0, 2.5^2), b_j ~ N(0, tau^2), y_i ~ Bernoulli(sigmoid(b0 + x_i'b))."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)

    def f(th):
        lt, b0, b = th[0], th[1], th[2:]
        eta = b0[None, :] + X @ b
        ll = np.sum(y[:, None] * eta - np.logaddexp(0.0, eta), axis=0)
        h = 0.5 * LOG2PI
        nb = b.shape[0]
        lp = -0.5 * lt * lt - h
        lp = lp + (-0.5 * (b0 / 2.5) ** 2 - math.log(2.5) - h)
        lp = lp + (-0.5 * np.sum(b * b, axis=0) * np.exp(-2 * lt) - nb * lt - nb * h)
        return lp + ll

    return f
