"""CPU tests (no GPU): pin the oracle against every fixture / known answer the reference's own
tests hold for the hot path, and the RNG / math contract against published vectors.

Each test names the reference test it mirrors (paths relative to the reference repo).
"""
import ctypes
import math

import numpy as np
import pytest
from scipy import stats

from oracle import pf_oracle as O
from oracle import psis as OP

# test/inverse_hessian.jl:19-20 — the reference's literal fixture
S0 = np.array([
    [1.719573, 3.294037, -2.008877, 3.901275, 0.214324, 0.400382, 0.113598, -1.804262, 0.465563, 2.465748],
    [-0.445476, -0.514915, 1.069617, -1.505506, -0.036535, -0.11386, -0.029737, 0.579662, -0.118694, -1.726258],
    [-0.013354, 0.1981, 0.081886, -0.194172, 0.010499, -0.007944, -0.001039, 0.061604, -0.002621, 0.11145],
    [0.00648, 0.255961, -0.011901, -0.059042, 0.013587, -0.002064, 0.000302, 0.023457, 0.00257, -0.002343],
    [-0.016408, 0.015005, 0.009654, 0.016735, -0.000245, -0.002825, -0.00106, 0.004272, -0.004578, 0.002275],
]).T
Y0 = -np.array([
    [-2.357935, -3.343312, 4.659008, -7.065433, -0.228045, -0.584164, -0.156837, 2.863289, -0.631753, -7.152021],
    [0.610851, 0.522617, -2.480668, 2.726559, 0.038874, 0.166123, 0.041055, -0.9199, 0.161064, 5.007094],
    [0.018312, -0.201064, -0.189911, 0.351657, -0.011171, 0.011591, 0.001434, -0.097763, 0.003557, -0.323265],
    [-0.008886, -0.25979, 0.027601, 0.106929, -0.014456, 0.003011, -0.000418, -0.037225, -0.003488, 0.006795],
    [0.022499, -0.015229, -0.02239, -0.030308, 0.00026, 0.004122, 0.001463, -0.00678, 0.006212, -0.006598],
]).T


def explicit_inverse_hessian(alpha, S, Y):
    """lbfgs_inverse_hessian_explicit, test/inverse_hessian.jl:8-14."""
    H0 = np.diag(alpha)
    B = np.hstack([H0 @ Y, S])
    R = np.triu(S.T @ Y)
    E = np.diag(np.diag(R))
    Rinv = np.linalg.inv(R)
    J = S.shape[1]
    D = np.block([[np.zeros((J, J)), -Rinv], [-Rinv.T, np.linalg.solve(R.T, E + Y.T @ H0 @ Y) @ Rinv]])
    return H0 + B @ D @ B.T


# ---- RNG / math contract --------------------------------------------------------------------
def test_philox4x32_10_known_answers():
    """Random123 kat_vectors (Salmon et al., SC'11) for philox4x32-10."""
    lib = O.clib()
    lib.pfo_philox.argtypes = [ctypes.c_uint32] * 6 + [ctypes.c_void_p]
    out = np.zeros(2, dtype=np.uint64)

    def call(c, k):
        lib.pfo_philox(*c, *k, out.ctypes.data)
        a, b = int(out[0]), int(out[1])
        return [a & 0xFFFFFFFF, a >> 32, b & 0xFFFFFFFF, b >> 32]

    assert call([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert call([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert call([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_pf_math_against_libm():
    """pf_math.h stays within 2 ulp of libm on the ranges the PSIS stage uses."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 20000), rng.normal(size=20000), [0.0, -0.0, 1e-300, -745.0, 709.0]])
    assert _ulps(O.pf_exp(x), np.exp(x)).max() <= 2
    y = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2, 20000), [5e-324, 1.0]])
    assert _ulps(O.pf_log(y), np.log(y))[np.log(y) != 0].max() <= 2
    z = np.concatenate([rng.uniform(-0.999, 10, 20000), rng.normal(size=20000) * 1e-8])
    assert _ulps(O.pf_log1p(z), np.log1p(z))[z != 0].max() <= 4
    assert _ulps(O.pf_expm1(z), np.expm1(z))[z != 0].max() <= 4
    assert np.isnan(O.pf_log(np.array([-1.0]))[0]) and O.pf_log(np.array([0.0]))[0] == -np.inf
    assert O.pf_exp(np.array([1000.0]))[0] == np.inf and O.pf_exp(np.array([-1000.0]))[0] == 0.0


def test_philox4x32_7_known_answers():
    """Random123 kat_vectors for philox4x32-7 (the round count of the normal stream, pf_rng.h), through
    the contract's own round function applied seven times; ten rounds of the same function give the
    philox4x32-10 vectors above."""
    lib = O.clib()
    out = np.zeros(4, dtype=np.uint32)

    def call(rounds, c, k):
        lib.pfo_philox_r(ctypes.c_int(rounds), *[ctypes.c_uint32(x) for x in c], *[ctypes.c_uint32(x) for x in k],
                         ctypes.c_void_p(out.ctypes.data))
        return [int(x) for x in out]

    assert call(7, [0, 0, 0, 0], [0, 0]) == [0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48]
    assert call(7, [0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x5207DDC2, 0x45165E59, 0x4D8EE751, 0x8C52F662]
    assert call(7, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0x4DFCCABA, 0x190A87F0, 0xC47362BA, 0xB6B5242A]
    assert call(10, [0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    # the contract call = philox4x32-7 with the fixed key and the documented counter layout
    seed, rp, stream, dp, cl = 0x0123456789ABCDEF, 1234, 2, 77, 5
    lib.pfo_bits4(ctypes.c_uint32(rp), ctypes.c_uint32(stream), ctypes.c_uint32(dp), ctypes.c_uint64(seed),
                  ctypes.c_uint32(cl), ctypes.c_void_p(out.ctypes.data))
    got = [int(x) for x in out]
    assert got == call(7, [rp | stream << 28, dp, ((seed & 0xFFFFFFFF) + cl) & 0xFFFFFFFF, seed >> 32],
                       [0xA4093822, 0x299F31D0])


def test_contract_normal_elementwise_definition():
    """pf_rng.h: element (row i, draw k) reads word 2 * ((k >> 3) & 1) + (i & 1) of the Philox call
    (row_pair = i >> 1, draw_pair = (k >> 4) * 8 + (k & 7)): the bulk generator, K3's per-lane 2 x 2
    patches and this element-wise definition are the same stream."""
    lib = O.clib()
    lib.pfo_normal_elem.restype = ctypes.c_double
    for seed, n, K in ((77, 37, 45), (2**63 + 5, 8, 16), (3, 1, 9)):
        U = np.asarray(O.contract_normals(seed, n, K))
        for i in range(n):
            for k in range(K):
                assert U[i, k] == lib.pfo_normal_elem(ctypes.c_uint64(seed), ctypes.c_uint32(i), ctypes.c_uint32(k))
    # a prefix in either direction is consistent (counter based): growing n or K never changes a variate
    assert np.array_equal(np.asarray(O.contract_normals(9, 40, 50))[:17, :23], np.asarray(O.contract_normals(9, 17, 23)))


def test_ziggurat_tables_are_consistent():
    """The packed 8-byte entries (layer edge with the accept threshold in its low 20 mantissa bits):
    thresholds are conservative, edges decrease, layers have equal areas to ~3e-7 (the packing
    perturbs an edge by < 2^-32 relative), f is exp(-x^2/2) at the packed edges."""
    lib = O.clib()
    xk = np.zeros(1024, dtype=np.uint64)
    f = np.zeros(1025)
    lib.pfo_zig_tables(ctypes.c_void_p(xk.ctypes.data), ctypes.c_void_p(f.ctypes.data))
    x = xk.view(np.float64)
    kq = (xk & np.uint64(0xFFFFF)).astype(np.int64)
    assert np.all(np.diff(x) < 0) and x[1] > 4.0 and x[0] > x[1]
    xn = np.append(x[1:], 0.0)
    assert np.all((kq - 1).clip(0) * 2.0**-20 * x < xn + 1e-300) and np.all((kq + 2) * 2.0**-20 * x > xn)
    assert np.allclose(f[:1024], np.exp(-0.5 * x * x), rtol=1e-15) and f[1024] == 1.0
    areas = x[1:] * (f[2:] - f[1:1024])
    v = x[1] * f[1] + math.sqrt(math.pi / 2) * math.erfc(x[1] / math.sqrt(2))
    assert np.abs(areas / v - 1).max() < 1e-6
    assert abs(x[0] * f[1] / v - 1) < 1e-6      # base strip: x_0 f(r) = v
    assert 0.995 < kq.sum() / 1024 / 2.0**20 < 0.9965   # fast-path share (99.57 %)


def _normal_stats(seed0, nseeds, n, K, thr, nbins):
    lib = O.clib()
    thr = np.asarray(thr, dtype=np.float64)
    edges = stats.norm.ppf(np.arange(1, nbins) / nbins)
    out = np.zeros(5 + thr.size + 1)
    hist = np.zeros(nbins, dtype=np.uint64)
    lib.pfo_normal_stats(ctypes.c_uint64(seed0), ctypes.c_int(nseeds), ctypes.c_int(n), ctypes.c_int(K),
                         ctypes.c_int(thr.size), ctypes.c_void_p(thr.ctypes.data), ctypes.c_int(nbins),
                         ctypes.c_void_p(edges.ctypes.data), ctypes.c_void_p(out.ctypes.data),
                         ctypes.c_void_p(hist.ctypes.data))
    return out, hist.astype(np.float64)


def test_contract_normals_are_standard_normal():
    """Moment / KS check of the engine's own normal stream (test/mvnormal.jl:71-107 checks the
    same of Julia's)."""
    u = np.asarray(O.contract_normals(2024, 1000, 1000)).ravel()
    n = u.size
    assert abs(u.mean()) < 5 / math.sqrt(n)
    assert abs(u.var() - 1) < 5 * math.sqrt(2 / n)
    assert abs(stats.kurtosis(u)) < 5 * math.sqrt(24 / n)
    assert stats.kstest(u[:200000], "norm").pvalue > 1e-4
    # tail beyond the ziggurat base strip edge has the right mass
    r = 4.038849846109505
    p = 2 * stats.norm.sf(r)
    big = np.asarray(O.contract_normals(7, 2000, 4000)).ravel()
    k = np.sum(np.abs(big) > r)
    assert abs(k - p * big.size) < 6 * math.sqrt(p * big.size)
    # seeds select independent streams; same seed reproduces
    v = np.asarray(O.contract_normals(2025, 1000, 100)).ravel()
    assert abs(np.corrcoef(u[:100000], v)[0, 1]) < 0.02
    assert np.array_equal(O.contract_normals(2024, 33, 7), O.contract_normals(2024, 33, 7))
    # rows of one draw and draws of one row are uncorrelated (the four words of a Philox call feed a
    # 2 x 2 patch: rows 2q, 2q+1 x draws k, k+8)
    U = np.asarray(O.contract_normals(11, 512, 4096))
    for a, b in ((U[0::2], U[1::2]), (U[:, :4088], U[:, 8:]), (U[:-1, :4088], U[1:, 8:])):
        c = float(np.mean(a * b))
        assert abs(c) < 5 / math.sqrt(a.size)


def test_contract_normals_large_sample():
    """1.3e8 variates (128 seeds x 1024 rows x 1000 draws — the shape of config 3's units) streamed
    through the contract in C: moments to 5 sigma, tail masses beyond 1, 2, 3, 4, r (the ziggurat base
    edge) and 5 to 5 sigma of their binomial spread, the slow-path share, and a 200-bin
    equal-probability chi-square."""
    thr = [1.0, 2.0, 3.0, 4.0, 4.038849846109505, 5.0]
    nbins = 200
    out, hist = _normal_stats(20261017, 128, 1024, 1000, thr, nbins)
    N = out[0]
    assert N == 128 * 1024 * 1000
    m1, m2, m3, m4 = out[1] / N, out[2] / N, out[3] / N, out[4] / N
    assert abs(m1) < 5 / math.sqrt(N)
    assert abs(m2 - 1) < 5 * math.sqrt(2 / N)
    assert abs(m3) < 5 * math.sqrt(15 / N)
    assert abs(m4 - 3) < 5 * math.sqrt(96 / N)
    for a, t in enumerate(thr):
        p = 2 * stats.norm.sf(t)
        assert abs(out[5 + a] - p * N) < 5 * math.sqrt(p * (1 - p) * N), (t, out[5 + a], p * N)
    slow = out[5 + len(thr)] / N
    assert 0.0040 < slow < 0.0046
    chi2 = float(np.sum((hist - N / nbins) ** 2 / (N / nbins)))
    assert stats.chi2.sf(chi2, nbins - 1) > 1e-4, chi2


# ---- inverse Hessian / Woodbury ---------------------------------------------------------------
def test_lbfgs_inverse_hessian_reference_fixture():
    """test/inverse_hessian.jl:16-44 (literal S0/Y0; J_eff in {0, 3, 5}; rotated ring buffer)."""
    rng = np.random.default_rng(1)
    N, J = S0.shape
    alpha = rng.random(N)
    B, D = O.lbfgs_inverse_hessian(alpha, S0, Y0, 0, 0)
    assert B.shape[1] == 0 or np.allclose(np.diag(alpha) + B @ D @ B.T, np.diag(alpha))
    B, D = O.lbfgs_inverse_hessian(alpha, S0, Y0, 3, 3)
    np.testing.assert_allclose(np.diag(alpha) + B @ D @ B.T, explicit_inverse_hessian(alpha, S0[:, :3], Y0[:, :3]),
                               rtol=1e-9, atol=1e-12)
    perm = [3, 4, 0, 1, 2]  # [4:J; 1:3] (1-based) — newest column (5) now sits at position 2
    S2, Y2 = S0[:, perm], Y0[:, perm]
    ilast = int(np.argmax(perm)) + 1
    B, D = O.lbfgs_inverse_hessian(alpha, S2, Y2, ilast, J)
    Hexp = explicit_inverse_hessian(alpha, S0, Y0)
    np.testing.assert_allclose(np.diag(alpha) + B @ D @ B.T, Hexp, rtol=1e-8, atol=1e-11)
    B, D = O.lbfgs_inverse_hessian(alpha, S0, Y0, J, J)
    np.testing.assert_allclose(np.diag(alpha) + B @ D @ B.T, Hexp, rtol=1e-8, atol=1e-11)


def test_inverse_hessians_reproduce_lbfgs_directions():
    """test/inverse_hessian.jl:46-76: with Hinit = y's/y'y the reconstructed H_l g_l is parallel to the
    step the optimiser actually took, and no update is rejected."""
    from scipy.optimize import minimize

    n, J = 10, 5
    rng = np.random.default_rng(3)
    A = rng.normal(size=(n, n))
    P = A @ A.T / n + np.eye(n)

    def f(x):
        return 0.5 * x @ P @ x + 0.05 * np.sum(x**4), P @ x + 0.2 * x**3

    pts, grads = [], []
    x0 = 3 * rng.normal(size=n)
    pts.append(x0.copy()); grads.append(-f(x0)[1])

    def cb(xk):
        pts.append(xk.copy()); grads.append(-f(xk)[1])

    minimize(f, x0, jac=True, method="L-BFGS-B", callback=cb, options=dict(maxcor=J, maxiter=25, gtol=1e-12))
    X, G = np.stack(pts, 1), np.stack(grads, 1)
    Hs, rejected, _ = O.lbfgs_inverse_hessians(X, G, history_length=J, hinit=O.nocedal_wright_scaling)
    assert rejected == 0
    for l in range(1, X.shape[1] - 1):
        step = X[:, l + 1] - X[:, l]
        p = Hs[l].mul(G[:, l])
        cos = step @ p / np.linalg.norm(step) / np.linalg.norm(p)
        assert cos > 1 - 1e-6, (l, cos)


def _rand_pd_diag(rng, n):
    return rng.random(n) + 0.05


@pytest.mark.parametrize("n,m", [(5, 8), (10, 8), (30, 12), (12, 12)])
def test_woodbury_identities_against_dense(n, m):
    """test/woodbury.jl:155-404 for A = Diagonal: dense reconstruction, logdet, mul, factor products,
    solves, invquad; right-factor structure [V 0; 0 I] Q' U (:26-36)."""
    rng = np.random.default_rng(n * 100 + m)
    alpha = _rand_pd_diag(rng, n)
    B = rng.normal(size=(n, m))
    Dh = rng.normal(size=(m, m))
    D = Dh @ Dh.T / m  # PSD => W is PD
    W = O.pdfactorize(alpha, np.asfortranarray(B), D)
    assert W.pd_ok
    Wmat = np.diag(alpha) + B @ D @ B.T
    np.testing.assert_allclose(W.dense(), Wmat)
    k = min(n, m)
    # R'R = W with R = diag(Vc, I) Q' U
    R = W.lmul_R(np.eye(n))
    np.testing.assert_allclose(R.T @ R, Wmat, rtol=1e-9, atol=1e-10)
    L = W.lmul_L(np.eye(n))
    np.testing.assert_allclose(L, R.T, rtol=1e-9, atol=1e-10)
    # structure of the right factor
    Q = W._q_apply(np.eye(n), trans=False)
    np.testing.assert_allclose(Q.T @ Q, np.eye(n), atol=1e-12)
    blk = np.eye(n)
    blk[:k, :k] = W.Vc
    np.testing.assert_allclose(R, blk @ Q.T @ np.diag(np.sqrt(alpha)), rtol=1e-9, atol=1e-10)
    # compact WY: Q = I - Vh T Vh'
    np.testing.assert_allclose(Q, np.eye(n) - W.Vh @ W.T[:k, :k] @ W.Vh.T, atol=1e-12)
    sign, ld = np.linalg.slogdet(Wmat)
    assert sign > 0 and abs(W.logdet() - ld) < 1e-9 * max(1, abs(ld))
    for shape in [(n,), (n, 3)]:
        x = rng.normal(size=shape)
        np.testing.assert_allclose(W.mul(x), Wmat @ x, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(W.ldiv_L(W.lmul_L(x)), x, rtol=1e-8, atol=1e-9)
        z = W.ldiv_L(x)  # whiten: dot(z, z) = x' W^-1 x   (test/woodbury.jl:311-355)
        ref = np.sum(x * np.linalg.solve(Wmat, x), axis=0)
        np.testing.assert_allclose(np.sum(z * z, axis=0), ref, rtol=1e-8)
    X = rng.normal(size=(n, 4))
    np.testing.assert_allclose(W.invquad(X), np.sum(X * np.linalg.solve(Wmat, X), axis=0), rtol=1e-8)


def test_fit_mvnormals_mean_and_covariances():
    """test/mvnormal.jl:8-29: Sigma_l == lbfgs_inverse_hessians output, mu = theta + Sigma g."""
    from tests.helpers import synthetic_trajectory

    X, G = synthetic_trajectory(10, 8, 5)
    mus, Hs, rej = O.fit_mvnormals(X, G, history_length=5)
    Hs2, rej2, _ = O.lbfgs_inverse_hessians(X, G, history_length=5)
    assert rej == rej2 and len(Hs) == X.shape[1]
    for l, (W, W2) in enumerate(zip(Hs, Hs2)):
        np.testing.assert_allclose(W.dense(), W2.dense())
        np.testing.assert_allclose(mus[:, l], X[:, l] + W.dense() @ G[:, l], rtol=1e-9, atol=1e-10)
    assert Hs[0].k == 0 and np.all(Hs[0].alpha == 1)  # H_0 = I (src/inverse_hessian.jl:38-40)


def test_rand_and_logpdf_matches_dense_mvn():
    """test/mvnormal.jl:31-68: x = mu + L u has logq == logpdf(MvNormal(mu, Sigma), x)."""
    from tests.helpers import synthetic_trajectory

    X, G = synthetic_trajectory(8, 6, 9)
    mus, Hs, _ = O.fit_mvnormals(X, G, history_length=4)
    u = np.random.default_rng(0).normal(size=(8, 50))
    for l in (0, 2, 6):
        x, logq = O.rand_and_logpdf(u, mus[:, l], Hs[l])
        Sig = Hs[l].dense()
        ref = stats.multivariate_normal(mus[:, l], Sig).logpdf(x.T)
        np.testing.assert_allclose(logq, ref, rtol=1e-9, atol=1e-9)
        # and the draws have the right covariance structure: L L' = Sigma
        L = Hs[l].lmul_L(np.eye(8))
        np.testing.assert_allclose(L @ L.T, Sig, rtol=1e-9, atol=1e-10)


# ---- ELBO ------------------------------------------------------------------------------------
def test_elbo_known_answer():
    """test/elbo.jl:7-28: target N(0, 0.08), fit N(0, sigma): ELBO = (1 - r^2)/2 + log r."""
    st = 0.08
    K = 200_000

    def logp(x):
        return stats.norm(0, st).logpdf(x[0])

    for i, sigma in enumerate([1e-3, 0.05, 0.8, 1.0, 5.0]):
        W = O.pdfactorize(np.array([sigma**2]), np.zeros((1, 0)), np.zeros((0, 0)))
        u = O.contract_normals(100 + i, 1, K)
        est = O.elbo_and_samples(u, logp, np.zeros(1), W)
        r = sigma / st
        assert abs(est["value"] - ((1 - r**2) / 2 + math.log(r))) < 4 * est["std_err"] + 1e-12
        np.testing.assert_array_equal(est["logr"], est["logp"] - est["logq"])
        assert np.isclose(est["value"], est["logr"].mean())
        assert np.isclose(est["std_err"], est["logr"].std(ddof=1) / math.sqrt(K))


def test_maximize_elbo_picks_matching_scale():
    """test/elbo.jl:30-54: among fits with sigma in [1e-3, .05, .08, 1, 5] the argmax is the third,
    ELBO ~ 0; reseeding reproduces; empty input -> (0, [])."""
    st = 0.08

    def logp(x):
        return stats.norm(0, st).logpdf(x[0])

    sig = [1e-3, 0.05, st, 1.0, 5.0]
    Hs = [None] + [O.pdfactorize(np.array([s**2]), np.zeros((1, 0)), np.zeros((0, 0))) for s in sig]
    mus = np.zeros((1, len(Hs)))
    seeds = np.arange(5, dtype=np.uint64) + 11
    lopt, ests = O.maximize_elbo(seeds, logp, mus, Hs, 100)
    assert lopt == 3 and abs(ests[2]["value"]) < 1e-12
    lopt2, ests2 = O.maximize_elbo(seeds, logp, mus, Hs, 100)
    assert lopt2 == lopt and all(np.array_equal(a["draws"], b["draws"]) for a, b in zip(ests, ests2))
    assert O.maximize_elbo(seeds[:0], logp, mus[:, :1], Hs[:1], 100) == (0, [])


def test_findmax_skipnan_table():
    """test/utils.jl:6-13."""
    nan = float("nan")
    assert O.findmax_skipnan([nan, 3.0, 1.0]) == (3.0, 2)
    v, i = O.findmax_skipnan([nan, nan])
    assert math.isnan(v) and i == 1
    assert O.findmax_skipnan([2.0, nan, 4.0]) == (4.0, 3)
    assert O.findmax_skipnan([1.0, 1.0]) == (1.0, 1)  # ties keep the earliest
    v, i = O.findmax_skipnan([])
    assert math.isnan(v) and i == 0
    assert not O.path_success(0, [], 0)
    assert not O.path_success(2, [dict(value=nan), dict(value=-math.inf)], 1)
    assert not O.path_success(2, [dict(value=nan), dict(value=-math.inf)], 2)
    assert O.path_success(2, [dict(value=nan), dict(value=-3.0)], 2)


# ---- PSIS / resample -------------------------------------------------------------------------
def test_log_importance_ratio_ordering():
    """test/resample.jl:62-89: ratios are draw-fastest, component-slowest."""
    from tests.helpers import synthetic_trajectory

    n, K_run, P = 4, 6, 3
    rng = np.random.default_rng(5)
    mus, Ws = [], []
    for p in range(P):
        X, G = synthetic_trajectory(n, 5, 40 + p)
        m, Hs, _ = O.fit_mvnormals(X, G, history_length=3)
        mus.append(m[:, -1]); Ws.append(Hs[-1])
    draws = rng.normal(size=(n, K_run, P))
    lr = O  # noqa
    ratios = OP.log_importance_ratios(O.logp_isonormal, mus, Ws, draws)
    for k in range(P):
        ref = stats.multivariate_normal(mus[k], Ws[k].dense())
        for j in range(K_run):
            x = draws[:, j, k]
            assert np.isclose(ratios[k * K_run + j], -0.5 * x @ x - ref.logpdf(x), rtol=1e-8, atol=1e-8)


def test_resample_membership_ids_and_degenerate_weights():
    """test/resample.jl:8-60."""
    dim, K_run, P, ndraws = 3, 10, 4, 20
    rng = np.random.default_rng(42)
    dpc = rng.normal(size=(dim, K_run, P))
    allc = dpc.reshape(dim, -1, order="F")
    draws, ids, inds = OP.resample(1, dpc, None, ndraws)  # uniform (psis_result === nothing)
    assert draws.shape == (dim, ndraws) and ids.shape == (ndraws,)
    assert ids.min() >= 1 and ids.max() <= P
    for c, cid in zip(draws.T, ids):
        assert any(np.array_equal(c, a) for a in allc.T)
        assert any(np.array_equal(c, a) for a in dpc[:, :, cid - 1].T)
    lw = np.full((K_run, P), -1000.0)
    lw[:, 0] = 0.0
    res = OP.psis(lw.reshape(-1, order="F"))
    draws, ids, _ = OP.resample(2, dpc, res, ndraws)
    assert np.all(ids == 1)
    for c in draws.T:
        assert any(np.array_equal(c, a) for a in dpc[:, :, 0].T)


def test_psis_result_properties_and_gpd_fit():
    """test/resample.jl:91-109 (length, sum(weights) ~ 1) + the Zhang-Stephens fit recovers the shape
    of a generalized-Pareto sample, and PSIS k-hat grows with the tail weight of the ratios."""
    rng = np.random.default_rng(7)
    lr = rng.normal(size=5000)
    r = OP.psis(lr)
    assert r["log_weights"].shape == (5000,) and abs(r["weights"].sum() - 1) < 1e-12
    assert r["tail_length"] == min(1000, math.ceil(3 * math.sqrt(5000)))
    for k_true in (-0.3, 0.2, 0.7):
        x = np.sort(stats.genpareto(k_true, scale=1.5).rvs(size=4000, random_state=rng))
        k, sigma = OP.fit_gpd(x)
        assert abs(k - k_true) < 0.08 and abs(sigma - 1.5) < 0.15
    k_light = OP.psis(rng.normal(size=20000) * 0.3)["pareto_k"]
    k_heavy = OP.psis(rng.standard_t(2, size=20000) * 2)["pareto_k"]
    assert k_light < 0.3 < 0.7 < k_heavy
    # too few draws for a tail fit: plain self-normalisation, k = NaN
    small = OP.psis(np.array([0.1, -0.2, 0.3, 0.0]))
    assert math.isnan(small["pareto_k"]) and abs(small["weights"].sum() - 1) < 1e-12


def test_resample_indices_follow_weights():
    w = np.array([0.5, 0.25, 0.125, 0.125])
    inds = OP.resample_indices(9, w, 4, 40000)
    freq = np.bincount(inds, minlength=5)[1:] / 40000
    assert np.abs(freq - w).max() < 0.01
    assert np.array_equal(inds, OP.resample_indices(9, w, 4, 40000))  # reproducible under reseed
    assert not np.array_equal(inds, OP.resample_indices(10, w, 4, 40000))


# ---- L-BFGS trajectory contract (row f1; src/optimize.jl:35-59, src/Pathfinder.jl:29-35) ----------
def test_lbfgs_isonormal_is_solved_in_one_newton_like_step():
    """test/singlepath.jl:13-41: on the iso-normal target the optimiser lands on the mode at once
    (history_length_effective == 1 there), the trace has the initial point plus few iterations."""
    from oracle import lbfgs as OL

    x0 = np.random.default_rng(0).uniform(-2, 2, size=10)
    X, FX, G, st, nev = OL.lbfgs_path(OL.FAMILY_ISONORMAL, x0)
    assert OL.STATUS[st] in ("gtol", "ftol")
    assert 2 <= X.shape[1] <= 4
    np.testing.assert_allclose(X[:, 0], x0)
    np.testing.assert_allclose(X[:, -1], 0.0, atol=1e-8)
    np.testing.assert_allclose(G, -X)                       # gradient of the LOG density (src/optimize.jl:96)
    np.testing.assert_allclose(FX, -0.5 * np.sum(X * X, axis=0))


def test_lbfgs_diag_normal_reaches_the_mean_and_matches_scipy():
    from scipy.optimize import minimize

    from oracle import lbfgs as OL

    rng = np.random.default_rng(1)
    n = 37
    mean, sd = rng.normal(size=n) * 3, rng.uniform(0.05, 20.0, size=n)
    x0 = rng.uniform(-2, 2, size=n)
    X, FX, G, st, nev = OL.lbfgs_path(OL.FAMILY_DIAGNORMAL, x0, mean=mean, sd=sd)
    assert OL.STATUS[st] in ("gtol", "ftol")
    np.testing.assert_allclose(X[:, -1], mean, atol=1e-5 * np.max(sd) ** 2)
    assert np.all(np.diff(FX) > 0)                          # monotone ascent of log p (Armijo)
    z = (X - mean[:, None]) / sd[:, None]
    np.testing.assert_allclose(G, -z / sd[:, None], rtol=1e-12, atol=1e-300)
    c0 = -np.sum(np.log(sd)) - 0.5 * n * np.log(2 * np.pi)
    np.testing.assert_allclose(FX, -0.5 * np.sum(z * z, axis=0) + c0, rtol=1e-12)
    ref = minimize(lambda x: (0.5 * np.sum(((x - mean) / sd) ** 2), (x - mean) / sd**2), x0, jac=True,
                   method="L-BFGS-B", options=dict(maxcor=6, gtol=1e-10, ftol=1e-15))
    assert abs((FX[-1] - c0) + ref.fun) < 1e-8
    assert X.shape[1] - 1 <= 3 * ref.nit + 10              # comparable iteration count


def test_lbfgs_funnel_trace_satisfies_wolfe_and_stops_on_caps():
    """Every recorded step satisfies the strong-Wolfe conditions the line search promises
    (=> s'y > 0, so lbfgs_inverse_hessians never has to reject an update for curvature)."""
    from oracle import lbfgs as OL

    n = 64
    x0 = np.random.default_rng(5).uniform(-10, 10, size=n)
    X, FX, G, st, nev = OL.lbfgs_path(OL.FAMILY_FUNNEL, x0, maxiters=40)
    assert X.shape[1] == 41 and OL.STATUS[st] == "maxiters"
    np.testing.assert_allclose(FX, O.logp_funnel(X), rtol=1e-12)
    S, Y = np.diff(X, axis=1), -np.diff(G, axis=1)
    assert np.all(np.sum(S * Y, axis=0) > 0)
    assert np.all(np.diff(FX) > 0)
    Xc, FXc, Gc, stc, _ = OL.lbfgs_path(OL.FAMILY_FUNNEL, x0, maxiters=1000, max_points=8)
    assert Xc.shape[1] == 8 and np.array_equal(Xc, X[:, :8]) and np.array_equal(Gc, G[:, :8])


def test_lbfgs_nonfinite_start_is_recorded_and_stops():
    """src/optimize.jl:94-105: the point is pushed, then the run stops."""
    from oracle import lbfgs as OL

    x0 = np.zeros(8)
    x0[0] = -800.0  # exp(800) = Inf
    x0[1:] = 1.0
    X, FX, G, st, nev = OL.lbfgs_path(OL.FAMILY_FUNNEL, x0)
    assert X.shape[1] == 1 and OL.STATUS[st] == "nonfinite" and nev == 1


# ---- resampling without replacement (test/resample.jl:31-34) --------------------------------------
def test_resample_without_replacement_unique_and_weighted():
    rng = np.random.default_rng(0)
    n, K_run, P = 3, 10, 4
    dpc = rng.normal(size=(n, K_run, P))
    draws, ids, inds = OP.resample(5, dpc, None, 5, replace=False)       # test/resample.jl:32-33
    assert len({tuple(c) for c in draws.T}) == 5 and len(set(inds)) == 5
    draws, ids, inds = OP.resample(5, dpc, None, K_run * P, replace=False)   # a permutation of the pool
    assert sorted(inds) == list(range(1, K_run * P + 1))
    assert np.array_equal(ids, -(-inds // K_run))
    with pytest.raises(ValueError):
        OP.resample(5, dpc, None, K_run * P + 1, replace=False)
    # weight only the first component (test/resample.jl:36-49): every draw comes from it
    lw = np.full((K_run, P), -1000.0)
    lw[:, 0] = 0.0
    ps = OP.psis(lw.reshape(-1, order="F"))
    _, ids, inds = OP.resample(9, dpc, ps, K_run, replace=False)
    assert np.all(ids == 1) and len(set(inds)) == K_run
    # inclusion frequencies follow the weights: P(first pick = i) = w_i
    w = np.array([0.5, 0.25, 0.125, 0.125])
    first = np.array([OP.resample_indices_norep(s, np.log(w), 4, 2)[0] for s in range(4000)])
    freq = np.bincount(first, minlength=5)[1:] / 4000.0
    assert np.max(np.abs(freq - w)) < 0.03


def test_lbfgs_dense_normal_and_hier_logistic_reach_the_scipy_optimum():
    """The GEMM-shaped families of the L-BFGS contract: gradients / log densities equal the host
    models', and the optimum equals SciPy's."""
    from scipy.optimize import minimize

    import pathfinder_b200 as pf
    from oracle import lbfgs as OL

    rng = np.random.default_rng(7)
    n = 30
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    prec = (Q / (rng.random(n) * 0.95 + 0.05)) @ Q.T
    prec = 0.5 * (prec + prec.T)
    mean = rng.normal(size=n)
    dm = pf.DenseNormal(mean, prec)
    X, FX, G, st, _ = OL.lbfgs_path(OL.FAMILY_DENSENORMAL, rng.uniform(-2, 2, n), 6, 1000, mean=mean, prec=prec)
    assert OL.STATUS[st] in ("gtol", "ftol")
    np.testing.assert_allclose(X[:, -1], mean, atol=1e-6)
    for l in (0, X.shape[1] // 2, X.shape[1] - 1):
        np.testing.assert_allclose(G[:, l], dm.grad(X[:, l]), rtol=1e-10, atol=1e-12)
        assert np.isclose(FX[l], dm.logp(X[:, l]), rtol=1e-10, atol=1e-12)

    nobs, p = 400, 10
    Xo = rng.normal(size=(nobs, p))
    yo = (rng.random(nobs) < 1.0 / (1.0 + np.exp(-(Xo @ (rng.normal(size=p) * 0.5))))).astype(np.float64)
    hm = pf.HierLogistic(Xo, yo)
    x0 = rng.uniform(-2, 2, p + 2)
    X, FX, G, st, _ = OL.lbfgs_path(OL.FAMILY_HLOGISTIC, x0, 6, 1000, Xobs=Xo, yobs=yo)
    assert OL.STATUS[st] in ("gtol", "ftol")
    for l in (0, 2, X.shape[1] - 1):
        np.testing.assert_allclose(G[:, l], hm.grad(X[:, l]), rtol=1e-9, atol=1e-9)
        assert np.isclose(FX[l], hm.logp(X[:, l]), rtol=1e-11)
    ref = minimize(lambda t: (-hm.logp(t), -hm.grad(t)), x0, jac=True, method="L-BFGS-B",
                   options=dict(maxcor=6, gtol=1e-10, ftol=1e-15, maxiter=2000))
    assert abs(FX[-1] + ref.fun) < 1e-6 * max(1.0, abs(ref.fun))
    assert np.max(np.abs(G[:, -1])) < 1e-4


def _psis_independent(log_ratios):
    """A second, independently written restatement of published PSIS (Vehtari et al., Algorithm 1;
    `gpdfit` of Zhang & Stephens 2009 as in the loo / ArviZ packages) in plain NumPy/libm — no
    shared code with oracle/psis.py — to cross-check the oracle's smoothed weights and k-hat."""
    x = np.array(log_ratios, dtype=np.float64)
    S = x.size
    M = int(min(math.ceil(S / 5), math.ceil(3 * math.sqrt(S))))
    mx = x.max()
    x = x - mx
    order = np.argsort(x, kind="stable")
    cutoff = x[order[S - M - 1]]
    tail = order[S - M:]                                # ascending
    xt = np.exp(x[tail]) - np.exp(cutoff)
    n = M
    m_est = 30 + int(math.sqrt(n))
    b = 1.0 - np.sqrt(m_est / (np.arange(1, m_est + 1) - 0.5))
    b = b / (3.0 * xt[int(n / 4 + 0.5) - 1]) + 1.0 / xt[-1]
    kk = np.log1p(-b[:, None] * xt[None, :]).mean(axis=1)
    L = n * (np.log(-(b / kk)) - kk - 1.0)
    with np.errstate(over="ignore"):
        w = 1.0 / np.exp(L[None, :] - L[:, None]).sum(axis=1)
    w = w / w.sum()
    b_post = float(np.sum(b * w))
    k_post = float(np.log1p(-b_post * xt).mean())
    sigma = -k_post / b_post
    k_hat = (n * k_post + 10 * 0.5) / (n + 10)
    p = (np.arange(n) + 0.5) / n
    q = sigma * np.expm1(-k_hat * np.log1p(-p)) / k_hat
    x[tail] = np.minimum(np.log(q + np.exp(cutoff)), 0.0)
    x = x - (np.log(np.sum(np.exp(x - x.max()))) + x.max())
    return x, k_hat, M


def test_psis_oracle_agrees_with_an_independent_restatement():
    rng = np.random.default_rng(12)
    for lr in (rng.normal(size=3000) * 2.0, rng.standard_t(3, size=8000) * 1.5, rng.gumbel(size=1200) - 50.0):
        got = OP.psis(lr)
        lw, k_hat, M = _psis_independent(lr)
        assert got["tail_length"] == M
        assert abs(got["pareto_k"] - k_hat) < 1e-10 * max(1.0, abs(k_hat))
        np.testing.assert_allclose(got["log_weights"], lw, rtol=1e-10, atol=1e-10)


def test_psis_and_resample_contract_properties():
    """Size-independent properties of the PSIS / resample contract: shift invariance and permutation
    equivariance of the smoothed weights, prefix consistency of both index streams (counter-based
    RNG: asking for more draws never changes the earlier ones)."""
    rng = np.random.default_rng(21)
    lr = rng.standard_t(4, size=2500) * 1.3
    base = OP.psis(lr)
    shifted = OP.psis(lr + 123.456)
    np.testing.assert_allclose(shifted["weights"], base["weights"], rtol=1e-9, atol=1e-300)
    assert abs(shifted["pareto_k"] - base["pareto_k"]) < 1e-9
    perm = rng.permutation(lr.size)
    permuted = OP.psis(lr[perm])
    np.testing.assert_allclose(permuted["weights"], base["weights"][perm], rtol=1e-9, atol=1e-300)
    w = base["weights"]
    long_ = OP.resample_indices(17, w, lr.size, 500)
    assert np.array_equal(OP.resample_indices(17, w, lr.size, 120), long_[:120])
    nr_all = OP.resample_indices_norep(17, base["log_weights"], lr.size, lr.size)
    assert np.array_equal(OP.resample_indices_norep(17, base["log_weights"], lr.size, 300), nr_all[:300])
    assert sorted(nr_all) == list(range(1, lr.size + 1))
    # heavier weights are picked earlier on average
    rank_of = np.empty(lr.size, dtype=np.int64)
    rank_of[nr_all - 1] = np.arange(lr.size)
    top = np.argsort(-w)[:50]
    assert rank_of[top].mean() < 0.25 * lr.size


def test_lbfgs_contract_agrees_with_an_independent_restatement():
    """pf_lbfgs.h is compiled into both K0 and the oracle, so their bit-exact agreement is one source under
    two compilers.  tests/indep_lbfgs.py states the same algorithm separately in plain NumPy (other
    summation order, libm exp): iterates, log densities, gradients, statuses and evaluation counts must
    agree — up to rounding, which grows along a funnel trajectory, hence short runs with a line search
    that brackets and zooms (more evaluations than iterations)."""
    from oracle import lbfgs as OL
    from tests import indep_lbfgs as IL

    def close(a, b, tol):
        assert a.shape == b.shape, (a.shape, b.shape)
        assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(a))), np.max(np.abs(a - b))

    zoomed = 0
    for seed in range(8):
        rng = np.random.default_rng(100 + seed)
        n, scale, iters = ((16, 2.0, 12), (32, 3.0, 14))[seed % 2]
        x0 = rng.uniform(-scale, scale, size=n)
        X, FX, G, st, nev = OL.lbfgs_path(OL.FAMILY_FUNNEL, x0, maxiters=iters)
        Xi, FXi, Gi, sti, nevi = IL.lbfgs_path(IL.density("funnel"), x0, maxiters=iters)
        assert OL.STATUS[st] == sti and nev == nevi, (seed, OL.STATUS[st], sti, nev, nevi)
        close(X, Xi, 1e-7); close(FX, FXi, 1e-7); close(G, Gi, 1e-6)
        zoomed += nev > X.shape[1] + 3
    assert zoomed >= 4  # the bracketing / zoom phase was exercised, not just first-trial acceptance

    rng = np.random.default_rng(7)
    x0 = rng.uniform(-2, 2, size=10)
    a, b = OL.lbfgs_path(OL.FAMILY_ISONORMAL, x0), IL.lbfgs_path(IL.density("isonormal"), x0)
    assert OL.STATUS[a[3]] == b[3] and a[4] == b[4]
    close(a[0], b[0], 1e-12); close(a[2], b[2], 1e-12)

    n = 24
    A = rng.normal(size=(n, n)) / np.sqrt(n)
    prec, mean = A @ A.T + np.diag(rng.uniform(0.5, 2.0, size=n)), rng.normal(size=n)
    x0 = rng.uniform(-2, 2, size=n)
    a = OL.lbfgs_path(OL.FAMILY_DENSENORMAL, x0, mean=mean, prec=prec)
    b = IL.lbfgs_path(IL.density("densenormal", mean=mean, prec=prec), x0)
    assert OL.STATUS[a[3]] == b[3] and a[4] == b[4]
    close(a[0], b[0], 1e-9); close(a[1], b[1], 1e-9); close(a[2], b[2], 1e-8)

    n = 37
    mean, sd = rng.normal(size=n) * 3, rng.uniform(0.05, 20.0, size=n)
    x0 = rng.uniform(-2, 2, size=n)
    a = OL.lbfgs_path(OL.FAMILY_DIAGNORMAL, x0, mean=mean, sd=sd, maxiters=30)
    b = IL.lbfgs_path(IL.density("diagnormal", mean=mean, sd=sd), x0, maxiters=30)
    assert a[4] == b[4]
    close(a[0], b[0], 1e-7); close(a[1], b[1], 1e-9)
    # both reach the same optimum when left to converge
    a = OL.lbfgs_path(OL.FAMILY_DIAGNORMAL, x0, mean=mean, sd=sd)
    b = IL.lbfgs_path(IL.density("diagnormal", mean=mean, sd=sd), x0)
    assert OL.STATUS[a[3]] == b[3] and abs(a[1][-1] - b[1][-1]) < 1e-8


def test_normal_contract_agrees_with_an_independent_restatement():
    """pf_rng.h is compiled into both the kernels and the oracle.  tests/indep_rng.py states the contract a
    second time in pure Python (integer Philox4x32-7, tables recomputed by the generator script, exact
    rational arithmetic for the fast path's single rounding, libm on the slow path): every variate of the
    C stream must be reproduced — bit for bit on the table-only path, to 1e-13 where exp / log enter."""
    from tests import indep_rng as IR

    lib = O.clib()
    lib.pfo_normal_elem.restype = ctypes.c_double
    total = slow = 0
    for seed, n, K in ((77, 120, 100), (2**63 + 5, 64, 80), (123456789012345, 33, 50)):
        U = np.asarray(O.contract_normals(seed, n, K))
        for i in range(n):
            for k in range(K):
                z, was_slow = IR.normal_elem(seed, i, k)
                total += 1
                slow += was_slow
                if was_slow:
                    assert abs(z - U[i, k]) <= 1e-13 * abs(U[i, k]), (seed, i, k, z, U[i, k])
                else:
                    assert z == U[i, k], (seed, i, k, z, U[i, k])
    assert 0.002 < slow / total < 0.008  # 0.43 % leave the fast path
    # the tail beyond r = 4.04 (5e-5 of the variates): find elements whose first word lands in the base
    # strip outside its core, and compare those too
    seed, tails = 1, 0
    for i in range(0, 400, 2):
        for dp in range(128):
            o = IR._call(i >> 1, 0, dp, seed, 0)
            for word in range(4):
                w = o[word]
                if (w >> 21) & 1023 == 0 and (w & 0xFFFFF) >= (IR._tables_cached()[0][0] & 0xFFFFF):
                    row, k = i + (word & 1), (dp >> 3) * 16 + (dp & 7) + 8 * (word >> 1)
                    z, was_slow = IR.normal_elem(seed, row, k)
                    ref = lib.pfo_normal_elem(ctypes.c_uint64(seed), ctypes.c_uint32(row), ctypes.c_uint32(k))
                    assert was_slow and abs(z) > 4.0 and abs(z - ref) <= 1e-13 * abs(ref), (row, k, z, ref)
                    tails += 1
    assert tails >= 2


def test_resample_index_stream_agrees_with_an_independent_restatement():
    """_resample with replacement (src/resample.jl:58-72) under the engine's contract, restated on Python
    integers: draw t takes 64 bits of Philox4x32-10 call t >> 1 on stream 3 (low pair for even t, high pair
    for odd t), target = floor(bits * Z / 2^64) on the fixed-point (2^52) cumulative weights, index = the
    first entry whose cumulative weight exceeds the target."""
    from tests import indep_rng as IR

    rng = np.random.default_rng(5)
    for seed, N, ndraws in ((9, 4, 300), (2**64 - 3, 257, 500), (123456789, 5000, 400)):
        w = rng.dirichlet(np.full(N, 0.3))
        fixed = [int(v * 4503599627370496.0) if v > 0.0 else 0 for v in w]
        cum, acc = [], 0
        for v in fixed:
            acc += v
            cum.append(acc)
        want = []
        for t in range(ndraws):
            q = t >> 1
            o = IR.philox4x32(10, ((q & 0x0FFFFFFF) | 3 << 28, q >> 28, seed & 0xFFFFFFFF, seed >> 32))
            bits = (o[2] | o[3] << 32) if t & 1 else (o[0] | o[1] << 32)
            target = (bits * acc) >> 64
            want.append(next(i for i, c in enumerate(cum) if c > target) + 1)
        assert np.array_equal(OP.resample_indices(seed, w, N, ndraws), want)
        uniform = [((lambda o, t: (o[2] | o[3] << 32) if t & 1 else (o[0] | o[1] << 32))(
            IR.philox4x32(10, (((t >> 1) & 0x0FFFFFFF) | 3 << 28, (t >> 1) >> 28, seed & 0xFFFFFFFF, seed >> 32)), t) * N >> 64) + 1
            for t in range(ndraws)]
        assert np.array_equal(OP.resample_indices(seed, None, N, ndraws), uniform)


def test_resample_without_replacement_agrees_with_an_independent_restatement():
    """replace = false (src/resample.jl:61-66) under the engine's contract, restated with Python integers
    and libm: key_i = log(E_i) - log w_i, E_i = -log(u_i), u_i = (top 53 bits of draw i's 64 + 0.5) 2^-53;
    the ndraws smallest keys in ascending order."""
    import math

    from tests import indep_rng as IR

    rng = np.random.default_rng(8)
    for seed, N, ndraws in ((4, 60, 60), (2**62 + 1, 700, 90)):
        lw = np.log(rng.dirichlet(np.full(N, 0.5)))
        keys = []
        for t in range(N):
            q = t >> 1
            o = IR.philox4x32(10, ((q & 0x0FFFFFFF) | 3 << 28, q >> 28, seed & 0xFFFFFFFF, seed >> 32))
            bits = (o[2] | o[3] << 32) if t & 1 else (o[0] | o[1] << 32)
            u = ((bits >> 11) + 0.5) * 2.0 ** -53
            keys.append(math.log(-math.log(u)) - lw[t])
        want = np.argsort(np.array(keys), kind="stable")[:ndraws] + 1
        got = OP.resample_indices_norep(seed, lw, N, ndraws)
        assert np.array_equal(got, want)
        assert len(set(got.tolist())) == ndraws
