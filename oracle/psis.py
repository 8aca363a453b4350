"""CPU oracle for Pareto-smoothed importance resampling  —  TEST INFRASTRUCTURE ONLY.

Restates, for the hot path's last stage (reference: src/resample.jl:58-95):

  * ``PSIS.psis(log_ratios)``  (src/resample.jl:78).  PSIS.jl (compat 0.2-0.9, Project.toml:64)
    is NOT in the reference tree and no Manifest pins it, so this follows the published
    algorithm: Vehtari, Simpson, Gelman, Yao, Gabry, "Pareto smoothed importance sampling"
    (JMLR 2024) with the generalized-Pareto fit of Zhang & Stephens (Technometrics 2009) and
    the weakly-informative shape prior, as also implemented by loo / ArviZ.
    PARITY UNPINNED: the reference tests check only ``sum(weights) ≈ 1``, length and the
    degenerate-weight behaviour (test/resample.jl:36-49, :103-108).
  * ``StatsBase.sample(rng, 1:N, pweights, ndraws; replace)``  (src/resample.jl:61-66).
    StatsBase (compat 0.33.17/0.34) is not in the tree either and its index stream is
    version dependent, so the engine fixes its own contract: inverse-CDF sampling on an
    integer (fixed-point, 2^52) cumulative weight table with 64 Philox bits per draw
    (pf_rng.h, stream 3).  PARITY UNPINNED against Julia; bit-exact against the GPU.

All floating-point work uses the bit-reproducible math contract (pf_math.h through
libpforacle.so) and the canonical summation orders below, so the GPU kernels can be compared
bit for bit on identical inputs.
"""
from __future__ import annotations

import math

import numpy as np

from . import pf_oracle as O


# ---- canonical summation orders (mirrored by csrc/psis kernels) ------------------------------
def _butterfly32(acc):
    """acc: (..., 32).  xor-butterfly 16, 8, 4, 2, 1; every lane ends with the same value."""
    idx = np.arange(32)
    for off in (16, 8, 4, 2, 1):
        acc = acc + acc[..., idx ^ off]
    return acc


def sum32(v):
    """Lane-strided sequential partial sums over 32 lanes, then the xor-butterfly."""
    v = np.asarray(v, dtype=np.float64)
    L = v.shape[-1]
    pad = (-L) % 32
    if pad:
        v = np.concatenate([v, np.zeros(v.shape[:-1] + (pad,))], axis=-1)
    rows = v.reshape(v.shape[:-1] + (-1, 32))
    acc = np.zeros(v.shape[:-1] + (32,))
    for r in range(rows.shape[-2]):
        acc = acc + rows[..., r, :]
    return _butterfly32(acc)[..., 0]


def sum1024(v):
    """Thread-strided (1024) sequential partials, butterfly inside each warp, butterfly over
    the 32 warp sums."""
    v = np.asarray(v, dtype=np.float64).ravel()
    pad = (-v.size) % 1024
    if pad:
        v = np.concatenate([v, np.zeros(pad)])
    rows = v.reshape(-1, 1024)
    acc = np.zeros(1024)
    for r in range(rows.shape[0]):
        acc = acc + rows[r]
    ws = _butterfly32(acc.reshape(32, 32))[:, 0]
    return _butterfly32(ws)[0]


def ordered_key(x):
    """Monotone map double -> uint64 (NaN sorts last, like Julia's isless)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = x.view(np.uint64)
    neg = (u >> np.uint64(63)) != 0
    key = np.where(neg, ~u, u | np.uint64(1 << 63))
    key = np.where(np.isnan(x), np.uint64(0xFFFFFFFFFFFFFFFF), key)
    return key


def tail_length(N):
    """PSIS.jl tail_length(r_eff=1, S): min(cld(S, 5), ceil(3 sqrt(S)))."""
    return min(-(-N // 5), int(math.ceil(3.0 * math.sqrt(N))))


def fit_gpd(x):
    """Zhang & Stephens (2009) empirical-Bayes GPD(mu=0) fit on the ascending sample x.
    Returns (k_post [unadjusted shape], sigma)."""
    M = x.size
    m = 30 + int(math.floor(math.sqrt(M)))
    xmax = x[M - 1]
    xq = x[int(math.floor(M / 4.0 + 0.5)) - 1]
    j = np.arange(1, m + 1, dtype=np.float64)
    with np.errstate(all="ignore"):
        b = 1.0 / xmax + (1.0 - np.sqrt(m / (j - 0.5))) / (3.0 * xq)
        k = sum32(O.pf_log1p(-(b[:, None] * x[None, :]))) / M
        ll = M * (O.pf_log(-(b / k)) - k - 1.0)
        w = 1.0 / sum32(O.pf_exp(ll[None, :] - ll[:, None]))
        w = w / sum32(w)
        b_post = sum32(b * w)
        k_post = sum32(O.pf_log1p(-(b_post * x))) / M
        sigma = -k_post / b_post
    return float(k_post), float(sigma)


def psis(log_ratios):
    """Returns dict(log_weights [normalised], weights, pareto_k, tail_length)."""
    logw = np.array(log_ratios, dtype=np.float64, copy=True).ravel()
    logw[np.isnan(logw)] = -np.inf  # engine contract: an undefined log ratio carries no weight
    N = logw.size
    M = tail_length(N)
    pareto_k = float("nan")
    with np.errstate(all="ignore"):
        if M >= 5:
            key = ordered_key(logw)
            order = np.lexsort((np.arange(N), key))  # ascending by (value, index)
            top = order[N - (M + 1):]
            icut, tail = top[0], top[1:]
            logu = logw[icut]
            tail_vals = logw[tail]
            if np.all(np.isfinite(tail_vals)):
                lw_max = tail_vals[M - 1]
                mu_s = float(O.pf_exp(np.array([logu - lw_max]))[0])
                x = O.pf_exp(tail_vals - lw_max) - mu_s
                k_post, sigma = fit_gpd(x)
                k_adj = (k_post * M + 5.0) / (M + 10.0)
                pareto_k = k_adj
                if math.isfinite(k_adj):
                    p = (np.arange(M, dtype=np.float64) + 0.5) / M
                    l1p = O.pf_log1p(-p)
                    if abs(k_adj) < 2.220446049250313e-16:
                        z = -l1p
                    else:
                        z = O.pf_expm1(-(k_adj * l1p)) / k_adj
                    q = sigma * z
                    val = O.pf_log(q + mu_s)
                    logw[tail] = np.minimum(val, 0.0) + lw_max
        finite_or_inf = logw[~np.isnan(logw)]
        mx = finite_or_inf.max() if finite_or_inf.size else float("nan")
        ssum = sum1024(O.pf_exp(logw - mx))
        lse = mx + float(O.pf_log(np.array([ssum]))[0])
        logw_n = logw - lse
        w = O.pf_exp(logw_n)
    return dict(log_weights=logw_n, weights=w, pareto_k=pareto_k, tail_length=M)


def weight_table(weights):
    """Fixed-point (2^52) integer weights and their inclusive cumulative sums (uint64)."""
    w = np.asarray(weights, dtype=np.float64)
    q = np.zeros(w.shape, dtype=np.uint64)
    ok = w > 0
    q[ok] = np.floor(w[ok] * 4503599627370496.0).astype(np.uint64)
    return q, np.cumsum(q, dtype=np.uint64)


def resample_indices(seed, weights, N, ndraws):
    """1-based indices, with replacement.  weights=None -> uniform (importance=false)."""
    bits = np.empty(ndraws, dtype=np.uint64)
    O.clib().pfo_resample_bits(int(seed) & (2**64 - 1), ndraws, bits.ctypes.data)
    mulhi = O.clib().pfo_mulhi64
    if weights is None:
        return np.array([mulhi(int(b), N) + 1 for b in bits], dtype=np.int64)
    _, cum = weight_table(weights)
    Z = int(cum[-1])
    if Z == 0:  # degenerate (all weights zero / NaN): engine falls back to uniform
        return np.array([mulhi(int(b), N) + 1 for b in bits], dtype=np.int64)
    targets = np.array([mulhi(int(b), Z) for b in bits], dtype=np.uint64)
    return np.searchsorted(cum, targets, side="right").astype(np.int64) + 1


def resample_indices_norep(seed, log_weights, N, ndraws):
    """1-based indices WITHOUT replacement (replace=false, src/resample.jl:61-66).  StatsBase's
    A-ExpJ runs on Julia's RNG stream (third party, parity unpinned); the engine contract is the
    same design as order statistics: key_i = log(E_i) - log w_i with E_i = -log(u_i) ~ Exp(1) from
    53 Philox bits of counter i; the ndraws smallest keys in ascending order, ties to the smaller
    index.  log_weights=None -> uniform (importance=false)."""
    if ndraws > N:
        raise ValueError("Cannot draw more samples without replacement.")
    bits = np.empty(N, dtype=np.uint64)
    O.clib().pfo_resample_bits(int(seed) & (2**64 - 1), N, bits.ctypes.data)
    u = ((bits >> np.uint64(11)).astype(np.float64) + 0.5) * 1.1102230246251565e-16
    e = -O.pf_log(u)
    lw = np.zeros(N) if log_weights is None else np.asarray(log_weights, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        key = np.where(np.isfinite(lw) | (lw == np.inf), O.pf_log(e) - lw, np.inf)
    key = np.where(np.isnan(lw), np.inf, key)
    # the device sorts the order-preserving integer image of the keys (NaN last), stable in the index
    b = key.view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    ordered = np.where(neg, ~b, b | np.uint64(1 << 63))
    ordered = np.where(np.isnan(key), np.uint64(0xFFFFFFFFFFFFFFFF), ordered)
    order = np.argsort(ordered, kind="stable")
    return order[:ndraws].astype(np.int64) + 1


def resample(seed, draws_per_component, psis_result, ndraws, replace=True):
    """reference: src/resample.jl:58-72.  draws_per_component: (n, K_run, P)."""
    n, K_run, P = draws_per_component.shape
    draws_all = draws_per_component.reshape(n, K_run * P, order="F")
    w = None if psis_result is None else psis_result["weights"]
    if replace:
        inds = resample_indices(seed, w, K_run * P, ndraws)
    else:
        lw = None if psis_result is None else psis_result["log_weights"]
        inds = resample_indices_norep(seed, lw, K_run * P, ndraws)
    draws = draws_all[:, inds - 1]
    ids = -(-inds // K_run)  # cld
    return draws, ids, inds


def log_importance_ratios(logp_fn, mus, Ws, draws_per_component):
    """reference: src/resample.jl:81-95 (draw-fastest, component-slowest)."""
    n, K_run, P = draws_per_component.shape
    out = np.empty((K_run, P))
    for k in range(P):
        x = draws_per_component[:, :, k]
        W = Ws[k]
        logq = -(n * O.LOG2PI + W.logdet()) / 2.0 - W.invquad(x - mus[k][:, None]) / 2.0
        out[:, k] = logp_fn(x) - logq
    return out.reshape(-1, order="F")
