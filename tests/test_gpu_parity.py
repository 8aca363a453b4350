"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU
oracle on the same seeded inputs.  Tolerances: ELBO and draws 1e-6 relative (north_star);
resample indices and PSIS weights bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def _engine(model, K, J=6, **kw):
    import pathfinder_b200 as pf

    return pf.Engine(model.n, model.family, model.blob, J, K, 0, **kw)


def _seeds(trajs, seed):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 2**64, size=X.shape[1] - 1, dtype=np.uint64) for X, _ in trajs]


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) if a.size else 0.0


def _sensitivity(model, trajs, seeds, K, J, orc):
    """Per-unit response of the ORACLE to a 1-ulp relative perturbation of its inputs.

    L-BFGS histories are often nearly collinear (cond(R_q) ~ 1e17 on funnel trajectories), so
    the reference algorithm itself is ill-conditioned there: two correct implementations (or
    the reference on two BLAS builds) differ by this much.  The parity tolerance is the
    north_star's 1e-6 on well-conditioned iterations and 50x this response elsewhere."""
    from tests.helpers import oracle_batch

    sens = []
    prng = np.random.default_rng(999)
    pert = []
    for X, G in trajs:
        pert.append((X * (1 + prng.choice([-1.0, 1.0], size=X.shape) * 1.2e-16),
                     G * (1 + prng.choice([-1.0, 1.0], size=G.shape) * 1.2e-16)))
    orc2 = oracle_batch(model, pert, seeds, K, J)
    for o, o2 in zip(orc, orc2):
        s = []
        for e, e2 in zip(o["ests"], o2["ests"]):
            with np.errstate(all="ignore"):
                d1 = abs(e["value"] - e2["value"]) / max(1.0, abs(e["value"]))
                d2 = np.nanmax(np.abs(e["draws"] - e2["draws"]) / np.maximum(1.0, np.abs(e["draws"])))
            s.append(np.nan_to_num(max(d1, d2), nan=0.0, posinf=1.0))
        sens.append(np.array(s))
    return sens


def _compare_batch(model, trajs, K, J, seed=1, normals=False, rtol=RTOL, cond_aware=False):
    import pathfinder_b200 as pf
    from tests.helpers import oracle_batch

    seeds = _seeds(trajs, seed)
    offsets, X, G = pf.Engine.pack(trajs)
    U = int(offsets[-1]) - len(trajs)
    nrm = None
    if normals:
        nrm = np.asfortranarray(np.random.default_rng(seed + 1).normal(size=(model.n, K, U)))
    eng = _engine(model, K, J, materialize_all=True)
    res = eng.elbo_batch(offsets, X, G, np.concatenate(seeds) if U else np.zeros(0, np.uint64), nrm,
                         draws=True, per_draw=True, fit=True, all_draws=True)
    orc = oracle_batch(model, trajs, seeds, K, J, normals=nrm)
    sens = _sensitivity(model, trajs, seeds, K, J, orc) if cond_aware else None
    worst_well_conditioned = 0.0
    n_units = n_relaxed = 0
    for p, o in enumerate(orc):
        sl = res.unit_slice(p)
        L = trajs[p][0].shape[1] - 1
        assert res.n_rejected[p] == o["rejected"]
        ev = np.array([e["value"] for e in o["ests"]])
        se = np.array([e["std_err"] for e in o["ests"]])
        tol = np.full(L, rtol) if sens is None else np.maximum(rtol, 50.0 * sens[p])
        for l in range(L):
            e = o["ests"][l]
            u = sl.start + l
            t = tol[l]
            assert abs(res.elbo[u] - ev[l]) <= t * max(1.0, abs(ev[l])) or (np.isnan(ev[l]) and np.isnan(res.elbo[u])), (p, l)
            assert abs(res.elbo_se[u] - se[l]) <= max(10 * t, 1e-5) * max(1e-4, abs(se[l])) or np.isnan(se[l]), (p, l)
            d = _rel(res.all_draws[:, :, u], e["draws"])
            assert d < t, (p, l, d, t)
            n_units += 1
            n_relaxed += int(t > rtol)
            if t == rtol:
                worst_well_conditioned = max(worst_well_conditioned, d)
            np.testing.assert_allclose(res.logq[:, u], e["logq"], rtol=t, atol=t)
            if t == rtol:
                np.testing.assert_allclose(res.logp[:, u], e["logp"], rtol=t, atol=t)
            else:
                # ill-conditioned iteration: the draws agree to t (relative); the funnel's logp
                # amplifies that by up to exp(-tau) |x|^2, so compare on the scale of the batch
                scale = max(1.0, float(np.nanmax(np.abs(e["logp"]))))
                np.testing.assert_allclose(res.logp[:, u], e["logp"], rtol=100 * t, atol=100 * t * scale)
        if res.best_iter[p] != o["lopt"]:
            # the argmax may legitimately flip between two iterations whose ELBOs agree within tol
            a, b = int(res.best_iter[p]) - 1, o["lopt"] - 1
            assert cond_aware and abs(ev[a] - ev[b]) <= 2 * max(tol[a], tol[b]) * max(1.0, abs(ev[b]))
            continue
        assert bool(res.success[p]) == o["success"]
        rtol_p = rtol if (sens is None or o["lopt"] == 0) else float(tol[o["lopt"] - 1])
        if o["lopt"] > 0:
            e = o["ests"][o["lopt"] - 1]
            assert _rel(res.draws[:, :, p], e["draws"]) < rtol_p
            np.testing.assert_allclose(res.draws_logp[:, p], e["logp"], rtol=100 * rtol_p, atol=rtol_p)
            np.testing.assert_allclose(res.draws_logq[:, p], e["logq"], rtol=rtol_p, atol=rtol_p)
            # the K5 re-materialisation must reproduce the ELBO-stage draws bit for bit
            ub = sl.start + o["lopt"] - 1
            assert np.array_equal(res.draws[:, :, p], res.all_draws[:, :, ub])
            assert np.array_equal(res.draws_logp[:, p], res.logp[:, ub])
            W = o["Hs"][o["lopt"]]
            np.testing.assert_allclose(res.fit["mu"][:, p], o["mus"][:, o["lopt"]], rtol=rtol_p, atol=rtol_p)
            np.testing.assert_allclose(res.fit["alpha"][:, p], W.alpha, rtol=1e-10)
            np.testing.assert_allclose(res.fit["logdet"][p], W.logdet(), rtol=1e-9, atol=1e-9)
            k = W.k
            if k and rtol_p == rtol:
                np.testing.assert_allclose(res.fit["vh"][:, :k, p], W.Vh[:, :k], rtol=1e-6, atol=1e-9)
                # LAPACK's geqrt blocks the compact-WY factor at 36 reflectors (Julia's qr: nb = min(k, 36)):
                # its T is [T1 | T2 ...] with T_b = the diagonal blocks of the full k x k factor exported here
                nb = W.T.shape[0]
                for b0 in range(0, k, nb):
                    b1 = min(k, b0 + nb)
                    np.testing.assert_allclose(res.fit["T"][p][b0:b1, b0:b1], W.T[: b1 - b0, b0:b1], rtol=1e-6,
                                               atol=1e-9)
                np.testing.assert_allclose(res.fit["Vc"][p][:k, :k], W.Vc[:k, :k], rtol=1e-6, atol=1e-9)
    eng.close()
    # how many units needed the relaxed (conditioning-aware) tolerance: counted, never silent
    RELAXED_LOG.append(dict(n=model.n, K=K, J=J, units=n_units, relaxed=n_relaxed))
    if not cond_aware:
        assert n_relaxed == 0
    return res, orc


RELAXED_LOG = []


def test_synthetic_small_device_rng():
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    model = pf.IsoNormal(16)
    trajs = [synthetic_trajectory(16, L, 10 + L) for L in (1, 3, 9)]
    _compare_batch(model, trajs, K=64, J=6)


def test_synthetic_host_normals():
    """Parity mode: normals supplied by the host (what a Julia caller does with its own RNG)."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    model = pf.IsoNormal(24)
    trajs = [synthetic_trajectory(24, L, 20 + L) for L in (2, 8)]
    _compare_batch(model, trajs, K=33, J=6, normals=True)


@pytest.mark.parametrize("n", [1, 5, 10, 13, 100, 300])
def test_dimensions_incl_n_smaller_than_2J(n):
    """dim in {1, 5, 10, 100} as test/singlepath.jl:13; n < 2J exercises min(n, 2J) reflectors."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    model = pf.IsoNormal(n)
    trajs = [synthetic_trajectory(n, L, 30 + n + L) for L in (0, 1, 4, 12)]
    _compare_batch(model, trajs, K=40, J=6)


def test_funnel_config2_trajectories():
    """BASELINE config 2 shape (100-dim funnel, history 6) on real L-BFGS trajectories."""
    import pathfinder_b200 as pf
    from tests.helpers import make_trajectories

    model = pf.Funnel(100)
    trajs = make_trajectories(model, 3, seed=5, init_scale=10, maxiters=60, min_len=5)
    _compare_batch(model, trajs, K=200, J=6, cond_aware=True)
    # the funnel's histories are rank deficient (tests/test_gpu_parity_fullsize.py): report, and bound, the share of
    # units whose draws needed the relaxed tolerance
    rec = RELAXED_LOG[-1]
    assert rec["units"] > 50 and rec["relaxed"] < rec["units"], rec
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in ("profiles", "gpurun_out"):
        path = os.path.join(root, d, "parity_report.json")
        if os.path.isdir(os.path.dirname(path)):
            try:
                rep = json.load(open(path))
            except Exception:
                rep = {}
            rep["cfg2_funnel100_k200_j6_three_paths"] = dict(units_compared=rec["units"], units_relaxed=rec["relaxed"],
                                                             strict_share=1.0 - rec["relaxed"] / rec["units"])
            try:
                json.dump(rep, open(path, "w"), indent=1, sort_keys=True)
            except OSError:
                pass


def test_history_10():
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    model = pf.IsoNormal(64)
    trajs = [synthetic_trajectory(64, 15, 77)]
    _compare_batch(model, trajs, K=32, J=10)


def test_rejected_updates_and_nonpd():
    """Negative-curvature steps are rejected and counted (src/inverse_hessian.jl:47,57)."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    n = 12
    X, G = synthetic_trajectory(n, 8, 3)
    G = G.copy()
    G[:, 4] = G[:, 3] - 2.0 * (G[:, 4] - G[:, 3])  # flip the curvature of one step
    model = pf.IsoNormal(n)
    res, orc = _compare_batch(model, [(X, G)], K=50, J=6)
    assert res.n_rejected[0] >= 1


def test_elbo_known_answer_diag_normal():
    """test/elbo.jl:8-28: target N(0, 0.08), fit N(0, sigma): ELBO = (1 - r^2)/2 + log r."""
    import pathfinder_b200 as pf

    st = 0.08
    model = pf.DiagNormal([0.0], [st])
    K = 200_000
    for sigma in (1e-3, 0.05, 0.8, 1.0, 5.0):
        # a 1-point-step trajectory whose fitted normal is N(0, sigma^2): theta1 = 0, grad1 = 0,
        # alpha = y's / y'y with s = -x0, y = g0 - g1 = -x0 / sigma^2  =>  alpha = sigma^2
        x0 = 1.0
        X = np.array([[x0, 0.0]])
        G = np.array([[-x0 / sigma**2, 0.0]])
        eng = _engine(model, K, 6)
        res = eng.elbo_batch(np.array([0, 2]), X, G, np.array([42], dtype=np.uint64), per_draw=True)
        r = sigma / st
        expect = (1 - r**2) / 2 + np.log(r)
        assert abs(res.elbo[0] - expect) < 4 * res.elbo_se[0] + 1e-12
        logr = res.logp[:, 0] - res.logq[:, 0]
        assert np.isclose(res.elbo[0], logr.mean(), rtol=1e-12)
        assert np.isclose(res.elbo_se[0], logr.std(ddof=1) / np.sqrt(K), rtol=1e-9)
        eng.close()


def test_psis_and_resample_bit_exact():
    from oracle import psis as OP
    import pathfinder_b200 as pf

    rng = np.random.default_rng(11)
    eng = _engine(pf.IsoNormal(3), 5)
    for N, K_run, scale in [(64000, 1000, 3.0), (8000, 1000, 1.0), (200, 20, 2.0), (60, 20, 0.5), (7, 7, 1.0)]:
        lr = rng.standard_t(4, size=N) * scale
        pool = np.asfortranarray(rng.normal(size=(3, N)))
        ref = OP.psis(lr)
        got = eng.psis_resample_host(lr, K_run, 1234, 500, True, pool=pool)
        assert got["tail_len"] == ref["tail_length"]
        assert np.array_equal(got["log_weights"], ref["log_weights"], equal_nan=True), N
        assert np.array_equal(got["weights"], ref["weights"], equal_nan=True)
        assert (got["pareto_k"] == ref["pareto_k"]) or (np.isnan(got["pareto_k"]) and np.isnan(ref["pareto_k"]))
        inds = OP.resample_indices(1234, ref["weights"], N, 500)
        assert np.array_equal(got["inds"], inds)
        assert np.array_equal(got["ids"], -(-inds // K_run))
        assert np.array_equal(got["draws"], pool[:, inds - 1])
        assert abs(got["weights"].sum() - 1) < 1e-12  # test/resample.jl:107
        # uniform resampling (psis_result === nothing)
        gu = eng.psis_resample_host(None, K_run, 99, 300, False, pool=pool)
        assert np.array_equal(gu["inds"], OP.resample_indices(99, None, N, 300))
    eng.close()


def test_psis_ties_and_degenerate_weights():
    """test/resample.jl:36-49: only the first component has weight -> all ids == 1."""
    from oracle import psis as OP
    import pathfinder_b200 as pf

    eng = _engine(pf.IsoNormal(3), 5)
    lw = np.full((10, 4), -1000.0)
    lw[:, 0] = 0.0
    lr = lw.reshape(-1, order="F")
    pool = np.asfortranarray(np.random.default_rng(0).normal(size=(3, 40)))
    got = eng.psis_resample_host(lr, 10, 42, 20, True, pool=pool)
    ref = OP.psis(lr)
    assert np.array_equal(got["weights"], ref["weights"])
    assert np.all(got["ids"] == 1)
    assert np.array_equal(got["inds"], OP.resample_indices(42, ref["weights"], 40, 20))
    # heavy ties inside the tail
    lr2 = np.round(np.random.default_rng(1).normal(size=5000), 1)
    got2 = eng.psis_resample_host(lr2, 50, 7, 100, True)
    ref2 = OP.psis(lr2)
    assert np.array_equal(got2["log_weights"], ref2["log_weights"], equal_nan=True)
    eng.close()


def test_multipathfinder_end_to_end_matches_oracle_pipeline():
    """multipathfinder on the funnel: pool order, log ratios, PSIS and indices vs the oracle."""
    from oracle import psis as OP
    import pathfinder_b200 as pf
    from tests.helpers import oracle_batch

    model = pf.Funnel(20)
    rng = np.random.default_rng(123)
    res = pf.multipathfinder(model, 50, nruns=4, ndraws_elbo=64, rng=rng, init_scale=5.0, maxiters=40)
    assert res.draws.shape == (20, 50)
    assert res.draw_component_ids.min() >= 1 and res.draw_component_ids.max() <= 4
    pool = np.concatenate([pr.draws for pr in res.pathfinder_results], axis=1)
    logr = np.concatenate([pr.draws_logp - pr.draws_logq for pr in res.pathfinder_results])
    ref = OP.psis(logr)
    assert np.array_equal(res.psis_result.weights, ref["weights"], equal_nan=True)
    assert np.array_equal(res.draws, pool[:, res.sample_inds - 1])
    assert np.array_equal(res.draw_component_ids, -(-res.sample_inds // 64))
    # every path's draws against the oracle with the same seeds is covered by _compare_batch;
    # here: the log ratios really are logp(x) - logq(x) of the returned draws
    from oracle import pf_oracle as O
    for pr in res.pathfinder_results:
        np.testing.assert_allclose(pr.draws_logp, O.logp_funnel(pr.draws), rtol=1e-9, atol=1e-9)


def _lean_vs_oracle(model, trajs, K, J, normals=False, rtol=RTOL):
    """The lean ELBO stage (single pass: log p from quadratic-form statistics, the normals generated
    once) against the oracle and against the generic two-pass kernel."""
    import pathfinder_b200 as pf
    from tests.helpers import oracle_batch

    seeds = _seeds(trajs, 3)
    offsets, X, G = pf.Engine.pack(trajs)
    U = int(offsets[-1]) - len(trajs)
    nrm = None
    if normals:
        nrm = np.asfortranarray(np.random.default_rng(5).normal(size=(model.n, K, U)))
    sd = np.concatenate(seeds) if U else np.zeros(0, np.uint64)
    orc = oracle_batch(model, trajs, seeds, K, J, normals=nrm)
    out = {}
    for two_pass in (False, True):
        eng = _engine(model, K, J, two_pass=two_pass)
        out[two_pass] = eng.elbo_batch(offsets, X, G, sd, nrm, draws=True, per_draw=True)
        eng.close()
    a, b = out[False], out[True]
    for p, o in enumerate(orc):
        sl = a.unit_slice(p)
        for l, e in enumerate(o["ests"]):
            u = sl.start + l
            scale = max(1.0, float(np.max(np.abs(e["logp"]))))
            np.testing.assert_allclose(a.logq[:, u], e["logq"], rtol=rtol, atol=rtol)
            np.testing.assert_allclose(a.logp[:, u], e["logp"], rtol=rtol, atol=rtol * scale)
            assert abs(a.elbo[u] - e["value"]) <= rtol * max(1.0, abs(e["value"]))
            # the two device formulations agree far below the parity tolerance
            np.testing.assert_allclose(a.logp[:, u], b.logp[:, u], rtol=1e-9, atol=1e-9 * scale)
            np.testing.assert_allclose(a.logq[:, u], b.logq[:, u], rtol=1e-13, atol=1e-12)  # |u|^2 rounding
        assert a.best_iter[p] == o["lopt"] == b.best_iter[p]
    assert np.array_equal(a.draws, b.draws)  # K5 is the same kernel in both engines


@pytest.mark.parametrize("n", [1, 7, 12, 16, 100, 257])
def test_lean_single_pass_isonormal(n):
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    _lean_vs_oracle(pf.IsoNormal(n), [synthetic_trajectory(n, L, 60 + n + L) for L in (1, 5, 11)], K=48, J=6)


def test_lean_single_pass_funnel_diagnormal_history10_host_normals():
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    rng = np.random.default_rng(8)
    n = 40
    trajs = [synthetic_trajectory(n, L, 80 + L, scale=0.3) for L in (2, 9)]
    _lean_vs_oracle(pf.Funnel(n), trajs, K=64, J=6)
    _lean_vs_oracle(pf.DiagNormal(rng.normal(size=n), rng.random(n) + 0.5), trajs, K=64, J=6)
    _lean_vs_oracle(pf.IsoNormal(n), [synthetic_trajectory(n, 14, 91)], K=40, J=10)
    _lean_vs_oracle(pf.IsoNormal(n), [synthetic_trajectory(n, 14, 92)], K=40, J=12)
    _lean_vs_oracle(pf.Funnel(n), trajs, K=24, J=6, normals=True)


def test_lean_single_pass_ring_mode_large_n():
    """n = 1500 does not fit the shared-memory stage: the factor record streams through the ring."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    n = 1500
    _lean_vs_oracle(pf.Funnel(n), [synthetic_trajectory(n, 4, 77, scale=0.2)], K=300, J=6)


def _rand_pd(rng, n):
    """rand_pd_mat of test/test_utils.jl:7-11: Q diag(U(0,1)) Q'."""
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    return Q @ np.diag(rng.random(n) + 0.05) @ Q.T


def test_dense_normal_config5_shape():
    """BASELINE config 5 at reduced size: correlated Gaussian (docs quickstart :17-24), history 10;
    log p goes through K8 (cuBLAS DGEMM + epilogue) on the materialised draws."""
    import pathfinder_b200 as pf
    from tests.helpers import make_trajectories

    rng = np.random.default_rng(6)
    n = 48
    Sigma = _rand_pd(rng, n)
    model = pf.DenseNormal(rng.normal(size=n), np.linalg.inv(Sigma))
    trajs = make_trajectories(model, 2, seed=7, init_scale=2.0, history_length=10, maxiters=30, min_len=4)
    _compare_batch(model, trajs, K=64, J=10, cond_aware=True)
    # lean mode (chunked K3 + K8, draws not kept) gives the same ELBO table
    seeds = _seeds(trajs, 1)
    offsets, X, G = pf.Engine.pack(trajs)
    eng = _engine(model, 64, 10)
    lean = eng.elbo_batch(offsets, X, G, np.concatenate(seeds), per_draw=True)
    eng.close()
    eng = _engine(model, 64, 10, materialize_all=True)
    full = eng.elbo_batch(offsets, X, G, np.concatenate(seeds), per_draw=True)
    eng.close()
    assert np.array_equal(lean.logp, full.logp) and np.array_equal(lean.elbo, full.elbo)


def test_hier_logistic_config4_shape():
    """BASELINE config 4 at reduced size: hierarchical logistic regression on synthetic X."""
    import pathfinder_b200 as pf
    from tests.helpers import make_trajectories

    rng = np.random.default_rng(4)
    p, nobs = 14, 96
    Xm = rng.normal(size=(nobs, p))
    beta = rng.normal(size=p) * 0.5
    y = (rng.random(nobs) < 1 / (1 + np.exp(-(0.3 + Xm @ beta)))).astype(np.float64)
    model = pf.HierLogistic(Xm, y)
    trajs = make_trajectories(model, 2, seed=9, init_scale=1.0, maxiters=40, min_len=4)
    _compare_batch(model, trajs, K=80, J=6, cond_aware=True)
    res = pf.multipathfinder(model, 30, nruns=3, ndraws_elbo=40, rng=np.random.default_rng(2), init_scale=1.0,
                             maxiters=40)
    assert res.draws.shape == (p + 2, 30)
    from oracle import pf_oracle as O
    f = O.make_logp_hier_logistic(Xm, y)
    for pr in res.pathfinder_results:
        np.testing.assert_allclose(pr.draws_logp, f(pr.draws), rtol=1e-9, atol=1e-9)


def test_topup_draws_and_resample_with_fresh_draws():
    """ndraws > ndraws_elbo tops the draws up from the fitted normal (src/singlepath.jl:228-230);
    resample(result; ndraws_per_run) redraws from the stored fits (src/resample.jl:102-109).
    Fresh draws must be x = mu + L u for the contract normals of their seed, with logq = logpdf."""
    from oracle import pf_oracle as O
    from oracle import psis as OP
    import pathfinder_b200 as pf

    model = pf.Funnel(12)
    rng = np.random.default_rng(77)
    res = pf.multipathfinder(model, 100, nruns=3, ndraws_elbo=16, rng=rng, init_scale=3.0, maxiters=30)
    # ndraws_per_run = max(16, cld(100, 3)) = 34 > 16: top-up happened
    assert all(pr.draws.shape == (12, 34) for pr in res.pathfinder_results)
    for pr in res.pathfinder_results:
        np.testing.assert_allclose(pr.draws_logp, O.logp_funnel(pr.draws), rtol=1e-9, atol=1e-9)
        mus, Hs, _ = O.fit_mvnormals(pr.optim_trace.points, pr.optim_trace.gradients, history_length=6)
        W, mu = Hs[pr.fit_iteration], mus[:, pr.fit_iteration]
        logq = -(12 * O.LOG2PI + W.logdet()) / 2.0 - W.invquad(pr.draws - mu[:, None]) / 2.0
        np.testing.assert_allclose(pr.draws_logq, logq, rtol=1e-6, atol=1e-6)
    pool = np.concatenate([pr.draws for pr in res.pathfinder_results], axis=1)
    logr = np.concatenate([pr.draws_logp - pr.draws_logq for pr in res.pathfinder_results])
    assert np.array_equal(res.psis_result.weights, OP.psis(logr)["weights"], equal_nan=True)
    assert np.array_equal(res.draws, pool[:, res.sample_inds - 1])
    assert np.array_equal(res.draw_component_ids, -(-res.sample_inds // 34))

    # resample with fresh draws: consistent with the stored fits, reproducible under reseed
    r2 = pf.resample(res, 40, rng=np.random.default_rng(5), ndraws_per_run=50)
    r3 = pf.resample(res, 40, rng=np.random.default_rng(5), ndraws_per_run=50)
    for j, pr in enumerate(r2.pathfinder_results):
        assert pr.draws.shape == (12, 50)
        assert np.array_equal(pr.draws, r3.pathfinder_results[j].draws)
        assert not np.array_equal(pr.draws[:, :16], res.pathfinder_results[j].draws[:, :16])
        mus, Hs, _ = O.fit_mvnormals(pr.optim_trace.points, pr.optim_trace.gradients, history_length=6)
        W, mu = Hs[pr.fit_iteration], mus[:, pr.fit_iteration]
        logq = -(12 * O.LOG2PI + W.logdet()) / 2.0 - W.invquad(pr.draws - mu[:, None]) / 2.0
        np.testing.assert_allclose(pr.draws_logq, logq, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(pr.draws_logp, O.logp_funnel(pr.draws), rtol=1e-9, atol=1e-9)
    assert np.array_equal(r2.sample_inds, r3.sample_inds)
    pool2 = np.concatenate([pr.draws for pr in r2.pathfinder_results], axis=1)
    assert r2.draws.shape == (12, 40) and np.array_equal(r2.draws, pool2[:, r2.sample_inds - 1])
    assert abs(r2.psis_result.weights.sum() - 1) < 1e-12
    # single path with ndraws > ndraws_elbo
    one = pf.pathfinder(model, ndraws_elbo=8, ndraws=20, rng=np.random.default_rng(1), init_scale=3.0, maxiters=30)
    assert one.draws.shape == (12, 20)
    np.testing.assert_allclose(one.draws_logp, O.logp_funnel(one.draws), rtol=1e-9, atol=1e-9)


def test_fresh_draws_match_oracle_on_a_well_conditioned_fit():
    """pfb_draw_from_fits: x = mu + L u for the contract normals of the given seed."""
    from oracle import pf_oracle as O
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    n, K = 20, 8
    trajs = [synthetic_trajectory(n, L, 300 + L) for L in (4, 9)]
    offsets, X, G = pf.Engine.pack(trajs)
    eng = _engine(pf.IsoNormal(n), K)
    eng.upload(offsets, X, G, np.zeros(13, dtype=np.uint64))
    best = [3, 9]
    eng.fit_only(best)
    seeds = np.array([11, 2**63 + 5], dtype=np.uint64)
    xd, lp, lq = eng.draw_from_fits(37, seeds)
    for p, (Xp, Gp) in enumerate(trajs):
        mus, Hs, _ = O.fit_mvnormals(Xp, Gp, history_length=6)
        x, logq = O.rand_and_logpdf(O.contract_normals(int(seeds[p]), n, 37), mus[:, best[p]], Hs[best[p]])
        assert _rel(xd[:, :, p], x) < RTOL
        np.testing.assert_allclose(lq[:, p], logq, rtol=RTOL, atol=RTOL)
        np.testing.assert_allclose(lp[:, p], O.logp_isonormal(x), rtol=RTOL, atol=RTOL)
    eng.close()


# ---- K0: device L-BFGS (row f1) ------------------------------------------------------------------
def _oracle_traces(model, x0s, J, maxiters, max_points=None, gtol=1e-8, ftol=1e-14):
    from oracle import lbfgs as OL

    kw = {}
    if model.family == OL.FAMILY_DIAGNORMAL:
        kw = dict(mean=model.mean, sd=model.sd)
    if model.family == OL.FAMILY_DENSENORMAL:
        kw = dict(mean=model.mean, prec=model.prec)
    if model.family == OL.FAMILY_HLOGISTIC:
        kw = dict(Xobs=model.X, yobs=model.y)
    return [OL.lbfgs_path(model.family, x0s[:, p], J, maxiters, max_points, gtol, ftol, **kw)
            for p in range(x0s.shape[1])]


@pytest.mark.parametrize("kind,n,J,scale,maxiters", [
    ("iso", 10, 6, 2.0, 1000), ("funnel", 5, 6, 3.0, 60), ("funnel", 100, 6, 10.0, 80),
    ("funnel", 1024, 6, 10.0, 40), ("funnel", 257, 10, 10.0, 50), ("diag", 333, 6, 2.0, 1000),
    ("dense", 150, 6, 2.0, 1000), ("hlogistic", 40, 6, 2.0, 300),
])
def test_device_lbfgs_trajectories_bit_exact(kind, n, J, scale, maxiters):
    """Kernel K0 against the CPU restatement of the same contract (pf_lbfgs.h): points, gradients,
    log densities, lengths, status and evaluation counts are bit-identical."""
    import pathfinder_b200 as pf

    rng = np.random.default_rng(n + J)
    if kind == "iso":
        model = pf.IsoNormal(n)
    elif kind == "funnel":
        model = pf.Funnel(n)
    elif kind == "hlogistic":
        nobs, p = 700, n - 2
        Xo = rng.normal(size=(nobs, p))
        yo = (rng.random(nobs) < 1.0 / (1.0 + np.exp(-(Xo @ (rng.normal(size=p) * 0.5))))).astype(np.float64)
        model = pf.HierLogistic(Xo, yo)
    elif kind == "dense":
        Sg = _rand_pd(rng, n)
        Pm = np.linalg.inv(Sg)
        model = pf.DenseNormal(rng.normal(size=n), 0.5 * (Pm + Pm.T))
    else:
        model = pf.DiagNormal(rng.normal(size=n) * 3, rng.uniform(0.05, 20.0, size=n))
    P = 7
    x0s = np.asfortranarray(rng.uniform(-scale, scale, size=(n, P)))
    eng = _engine(model, 16, J)
    npts, st, nev = eng.lbfgs_batch(x0s, maxiters)
    off, X, FX, G = eng.lbfgs_download()
    ref = _oracle_traces(model, x0s, J, maxiters)
    for p, (Xo, FXo, Go, sto, nevo) in enumerate(ref):
        assert npts[p] == Xo.shape[1] and st[p] == sto and nev[p] == nevo, (p, npts[p], Xo.shape[1], st[p], sto)
        sl = slice(off[p], off[p + 1])
        assert np.array_equal(X[:, sl], Xo, equal_nan=True)
        assert np.array_equal(G[:, sl], Go, equal_nan=True)
        assert np.array_equal(FX[sl], FXo, equal_nan=True)
    eng.close()


def test_device_lbfgs_independent_anchors_on_the_gpu():
    """K0 against anchors that do NOT share its source (the bit-exact oracle compiles the same pf_lbfgs.h):
    analytic optima of the Gaussian families, SciPy's L-BFGS-B optimum of the hierarchical logistic model,
    monotone log densities and recorded gradients equal to the model's own gradient (src/optimize.jl:94-101)."""
    import pathfinder_b200 as pf
    from scipy.optimize import minimize

    rng = np.random.default_rng(12)
    n = 60
    Sg = _rand_pd(rng, n)
    Pm = np.linalg.inv(Sg)
    mean = rng.normal(size=n)
    nobs = 300
    Xo = rng.normal(size=(nobs, 10))
    yo = (rng.random(nobs) < 1.0 / (1.0 + np.exp(-(Xo @ (rng.normal(size=10) * 0.7))))).astype(np.float64)
    cases = [(pf.IsoNormal(n), np.zeros(n)), (pf.DiagNormal(mean, rng.uniform(0.2, 5.0, size=n)), mean),
             (pf.DenseNormal(mean, 0.5 * (Pm + Pm.T)), mean), (pf.HierLogistic(Xo, yo), None)]
    for model, opt in cases:
        m = model.n
        x0s = np.asfortranarray(rng.uniform(-2, 2, size=(m, 4)))
        eng = _engine(model, 8, 6)
        npts, st, _ = eng.lbfgs_batch(x0s, 1000)
        off, X, FX, G = eng.lbfgs_download()
        if opt is None:
            r = minimize(lambda x: -model.logp(x), x0s[:, 0], jac=lambda x: -model.grad(x), method="L-BFGS-B",
                         options=dict(maxiter=2000, ftol=1e-15, gtol=1e-10))
            opt = r.x
        for p in range(4):
            sl = slice(off[p], off[p + 1])
            assert st[p] in (0, 1), st[p]                                   # a tolerance, not maxiters / failure
            assert np.all(np.diff(FX[sl]) >= -1e-9 * np.maximum(1.0, np.abs(FX[sl][:-1])))   # monotone ascent
            xL = X[:, sl][:, -1]
            assert np.max(np.abs(xL - opt)) < 2e-4 * max(1.0, np.max(np.abs(opt))), np.max(np.abs(xL - opt))
            for l in (0, (off[p + 1] - off[p]) // 2, off[p + 1] - off[p] - 1):
                np.testing.assert_allclose(G[:, sl][:, l], model.grad(X[:, sl][:, l]), rtol=1e-9, atol=1e-9)
                np.testing.assert_allclose(FX[sl][l], model.logp(X[:, sl][:, l]), rtol=1e-10, atol=1e-9)
        eng.close()


def test_device_lbfgs_capacity_nonfinite_and_unsupported_family():
    import pathfinder_b200 as pf

    n = 16
    model = pf.Funnel(n)
    x0s = np.asfortranarray(np.random.default_rng(3).uniform(-5, 5, size=(n, 3)))
    x0s[0, 1], x0s[1:, 1] = -800.0, 1.0  # exp(800) = Inf: recorded, then the run stops (src/optimize.jl:103-105)
    eng = _engine(model, 8)
    npts, st, nev = eng.lbfgs_batch(x0s, 1000, max_points=9)
    assert list(npts) == [9, 1, 9] and st[1] == 4 and st[0] == 2
    ref = _oracle_traces(model, x0s, 6, 1000, 9)
    off, X, FX, G = eng.lbfgs_download()
    assert np.array_equal(X[:, :9], ref[0][0]) and np.array_equal(G[:, 10:], ref[2][2])
    # the failed path has L = 0: no ELBO units, success = False, the others run normally
    seeds = np.arange(16, dtype=np.uint64) + 1
    eng.batch_from_lbfgs(seeds)
    eng.run()
    res = eng.download()
    assert list(res.success) == [True, False, True] and res.best_iter[1] == 0
    eng.close()
    hm = pf.HostModel(4, lambda X: -0.5 * (X * X).sum(axis=0), lambda x: -x)   # host closures: no device L-BFGS
    eng = pf.Engine.for_model(hm, 6, 8, 0)
    with pytest.raises(pf.PfbError) as ei:
        eng.lbfgs_batch(np.zeros((4, 2)), 10)
    assert ei.value.code == -3
    eng.close()


def test_device_optimizer_pipeline_equals_host_fed_pipeline():
    """multipathfinder(optimizer='device') == the engine fed, through the host upload path, with the
    oracle's trajectories for the same inits: K0's traces never leave the device, yet every
    downstream number (ELBO table, best iterations, PSIS weights, indices, draws) is identical."""
    import pathfinder_b200 as pf
    from pathfinder_b200.api import _draw_seeds, _uniform_init

    n, P, K, ndraws, maxiters = 48, 5, 64, 100, 30
    model = pf.Funnel(n)
    r = pf.multipathfinder(model, ndraws, nruns=P, ndraws_elbo=K, rng=np.random.default_rng(11), init_scale=4.0,
                           maxiters=maxiters, optimizer="device")
    # replay: same rng protocol on the host
    rng = np.random.default_rng(11)
    run_seeds = _draw_seeds(rng, P)
    prngs = [np.random.Generator(np.random.Philox(key=int(s))) for s in run_seeds]
    inits = np.stack([_uniform_init(prngs[p], n, 4.0) for p in range(P)], axis=1)
    seed = int(_draw_seeds(rng, 1)[0])
    ref = _oracle_traces(model, inits, 6, maxiters)
    seeds = [_draw_seeds(prngs[p], ref[p][0].shape[1] - 1) for p in range(P)]
    eng = _engine(model, K)
    offsets, X, G = pf.Engine.pack([(t[0], t[2]) for t in ref])
    res = eng.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=True, fit=True)
    rr = eng.psis_resample(seed, ndraws, True)
    assert all(pr.num_tries == 1 for pr in r.pathfinder_results)
    for p, pr in enumerate(r.pathfinder_results):
        assert np.array_equal(pr.optim_trace.points, ref[p][0])
        assert np.array_equal(pr.optim_trace.log_densities, ref[p][1])
        assert np.array_equal([e.value for e in pr.elbo_estimates], res.elbo[res.unit_slice(p)], equal_nan=True)
        assert pr.fit_iteration == res.best_iter[p]
    assert np.array_equal(r.sample_inds, rr["inds"]) and np.array_equal(r.draws, rr["draws"])
    assert np.array_equal(r.psis_result.weights, rr["weights"])
    eng.close()


def test_resample_without_replacement_bit_exact():
    """K7b (replace=false, src/resample.jl:61-66) against the oracle: identical indices, ids and
    gathered draws, weighted and uniform; uniqueness (test/resample.jl:31-34); ndraws > N errors."""
    import pathfinder_b200 as pf
    from oracle import psis as OP

    rng = np.random.default_rng(21)
    n, K_run, P = 7, 50, 9
    N = K_run * P
    pool = np.asfortranarray(rng.normal(size=(n, N)))
    logr = rng.normal(size=N) * 3.0
    logr[5] = -np.inf
    eng = _engine(pf.IsoNormal(n), K_run)
    for nd in (1, 40, N):
        r = eng.psis_resample_host(logr, K_run, 77, nd, True, pool=pool, replace=False)
        ps = OP.psis(logr)
        assert np.array_equal(r["log_weights"], ps["log_weights"], equal_nan=True)
        ref = OP.resample_indices_norep(77, ps["log_weights"], N, nd)
        assert np.array_equal(r["inds"], ref)
        assert len(set(r["inds"])) == nd
        assert np.array_equal(r["ids"], -(-ref // K_run))
        assert np.array_equal(r["draws"], pool[:, ref - 1])
        ru = eng.psis_resample_host(None, K_run, 78, nd, False, pool=pool, replace=False)
        assert np.array_equal(ru["inds"], OP.resample_indices_norep(78, None, N, nd))
    assert r["inds"][-1] == 6  # the zero-weight entry is taken last
    with pytest.raises(pf.PfbError) as ei:
        eng.psis_resample_host(logr, K_run, 1, N + 1, True, pool=pool, replace=False)
    assert ei.value.code == -1
    eng.close()


# ---- row f2: host-callback target density ----------------------------------------------------------
def test_host_callback_model_matches_device_family_and_oracle():
    """The same funnel, once as the registered device family and once as a HOST closure evaluated
    through the pinned-memory pipeline (several chunks in flight): ELBO tables, best iterations,
    pool log ratios agree; and the host-callback path agrees with the oracle."""
    import pathfinder_b200 as pf
    from oracle import pf_oracle as O
    from tests.helpers import make_trajectories, oracle_batch

    n, K, J = 40, 256, 6
    dev_model = pf.Funnel(n)
    calls = []

    def logp_batch(X):
        calls.append(X.shape)
        return O.logp_funnel(X)

    host_model = pf.HostModel(n, logp_batch, dev_model.grad)
    trajs = make_trajectories(dev_model, 4, seed=5, init_scale=3.0, maxiters=25, min_len=4)
    seeds = _seeds(trajs, 9)
    offsets, X, G = pf.Engine.pack(trajs)
    e1 = _engine(dev_model, K, J, two_pass=True)
    r1 = e1.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=True, per_draw=True)
    e2 = pf.Engine.for_model(host_model, J, K, 0)
    r2 = e2.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=True, per_draw=True)
    assert sum(s[1] for s in calls) == K * r1.elbo.size and all(s[0] == n for s in calls)
    assert np.array_equal(r1.logq, r2.logq, equal_nan=True)          # same sampling kernel
    assert np.array_equal(r1.draws, r2.draws, equal_nan=True)
    ok = np.isfinite(r1.elbo)
    assert np.array_equal(ok, np.isfinite(r2.elbo))
    assert _rel(r2.elbo[ok], r1.elbo[ok]) < 1e-9                     # exp() of libm vs libdevice
    assert np.array_equal(r1.best_iter, r2.best_iter)
    assert _rel(r2.draws_logp, r1.draws_logp) < 1e-9
    # against the oracle (same tolerance policy as the device families)
    orc = oracle_batch(host_model, trajs, seeds, K, J)
    sens = _sensitivity(host_model, trajs, seeds, K, J, orc)
    u = 0
    for p, o in enumerate(orc):
        for l, est in enumerate(o["ests"]):
            if np.isfinite(est["value"]):
                tol = max(RTOL, 50 * sens[p][l]) * max(1.0, abs(est["value"]))
                assert abs(r2.elbo[u] - est["value"]) <= tol, (p, l)
            u += 1
    # PSIS + resampling on the pool, and fresh draws through the callback
    rr = e2.psis_resample(3, 50, True)
    assert rr["draws"].shape == (n, 50)
    xd, lp, lq = e2.draw_from_fits(32, np.arange(4, dtype=np.uint64) + 5)
    fin = np.isfinite(lp)
    np.testing.assert_allclose(lp[fin], O.logp_funnel(xd.reshape(n, -1, order="F")).reshape(32, 4, order="F")[fin],
                               rtol=1e-12)
    e1.close(); e2.close()


def test_host_callback_exception_is_raised_not_swallowed():
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    def bad(X):
        raise RuntimeError("boom")

    n = 6
    model = pf.HostModel(n, bad, lambda x: -x)
    X, G = synthetic_trajectory(n, 5, 1)
    offsets, Xp, Gp = pf.Engine.pack([(X, G)])
    eng = pf.Engine.for_model(model, 6, 16, 0)
    with pytest.raises(RuntimeError, match="boom"):
        eng.elbo_batch(offsets, Xp, Gp, np.arange(5, dtype=np.uint64))
    eng.close()


def test_pathfinder_api_with_a_host_model():
    """pathfinder() on a correlated Gaussian given only as host closures (the reference's generic
    entry point, src/singlepath.jl:142-152): recovers mean and covariance like test/multipath.jl:12-61."""
    import pathfinder_b200 as pf

    rng = np.random.default_rng(3)
    n = 10
    A = rng.normal(size=(n, n))
    Sigma = A @ A.T / n + np.eye(n) * 0.5
    Pm = np.linalg.inv(Sigma)
    mu = rng.normal(size=n)
    model = pf.HostModel(n, lambda X: -0.5 * np.einsum("ij,ij->j", X - mu[:, None], Pm @ (X - mu[:, None])),
                         lambda x: -(Pm @ (x - mu)))
    r = pf.multipathfinder(model, 4000, nruns=8, ndraws_elbo=200, rng=np.random.default_rng(0))
    assert r.draws.shape == (n, 4000)
    assert np.max(np.abs(r.draws.mean(axis=1) - mu)) < 0.15
    assert np.max(np.abs(np.cov(r.draws) - Sigma)) < 0.35


def test_lazy_per_path_draws_on_a_caller_owned_engine():
    """With a caller-owned engine multipathfinder leaves PathfinderResult.draws on the device; they
    are fetched on first access, or automatically before the engine's pool is overwritten — and
    equal the eagerly downloaded ones bit for bit."""
    import pathfinder_b200 as pf

    n, P, K = 24, 6, 48
    model = pf.Funnel(n)
    kw = dict(nruns=P, ndraws_elbo=K, init_scale=3.0, maxiters=20, optimizer="device", ntries=1)
    eager = pf.multipathfinder(model, 30, rng=np.random.default_rng(4), **kw)          # engine owned by the call
    eng = pf.Engine.for_model(model, 6, K, 0)
    lazy1 = pf.multipathfinder(model, 30, rng=np.random.default_rng(4), engine=eng, **kw)
    assert all(pr._lazy is not None for pr in lazy1.pathfinder_results)
    assert np.array_equal(lazy1.draws, eager.draws) and np.array_equal(lazy1.sample_inds, eager.sample_inds)
    # access path 2 only -> one fetch serves the whole batch
    assert np.array_equal(lazy1.pathfinder_results[2].draws, eager.pathfinder_results[2].draws)
    lazy2 = pf.multipathfinder(model, 30, rng=np.random.default_rng(4), engine=eng, **kw)
    # a third run overwrites the pool: lazy2's draws are handed over first
    lazy3 = pf.multipathfinder(model, 30, rng=np.random.default_rng(5), engine=eng, **kw)
    for a, b in zip(lazy2.pathfinder_results, eager.pathfinder_results):
        assert np.array_equal(a.draws, b.draws, equal_nan=True)
        assert np.array_equal(a.draws_logp, b.draws_logp, equal_nan=True)
    eng.close()   # closing hands over lazy3's draws
    assert lazy3.pathfinder_results[0].draws.shape == (n, K)


def test_edge_cases_single_draw_and_empty_batch():
    """K = 1: mean = the single log ratio, std_err = NaN (var with K - 1 = 0, src/elbo.jl:18);
    P = 0 and all-paths-L=0 batches run and return empty / failed results instead of erroring."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    model = pf.IsoNormal(6)
    X, G = synthetic_trajectory(6, 3, 8)
    eng = _engine(model, 1)
    offsets, Xp, Gp = pf.Engine.pack([(X, G)])
    res = eng.elbo_batch(offsets, Xp, Gp, np.arange(3, dtype=np.uint64), per_draw=True)
    assert np.allclose(res.elbo, (res.logp - res.logq)[0]) and np.all(np.isnan(res.elbo_se))
    assert res.best_iter[0] >= 1 and res.success[0]
    # every path has only its initial point: no units, nothing succeeds (src/singlepath.jl:309-314)
    res0 = eng.elbo_batch(np.array([0, 1, 2]), Xp[:, :2], Gp[:, :2], np.zeros(0, np.uint64))
    assert res0.elbo.size == 0 and list(res0.best_iter) == [0, 0] and not res0.success.any()
    # their draws come from the identity fit of iteration 0, N(theta_0 + grad_0, I) (src/singlepath.jl:224-228)
    assert np.all(np.isfinite(res0.draws)) and res0.draws.shape == (6, 1, 2)
    assert np.all(np.abs(res0.draws[:, 0, :] - (Xp[:, :2] + Gp[:, :2])) < 8.0)
    # no paths at all
    resE = eng.elbo_batch(np.array([0]), np.zeros((6, 0), order="F"), np.zeros((6, 0), order="F"),
                          np.zeros(0, np.uint64))
    assert resE.elbo.size == 0 and resE.best_iter.size == 0
    eng.close()


def test_reused_pinned_output_buffers_give_identical_results():
    """download(into=...) / psis_resample(into=...) write into the caller's page-locked arrays of an
    earlier call (pfb_host_register); the numbers are the same as with fresh arrays."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    n, K = 20, 64
    model = pf.IsoNormal(n)
    trajs = [synthetic_trajectory(n, L, 40 + L) for L in (3, 7)]
    seeds = np.concatenate(_seeds(trajs, 2))
    offsets, X, G = pf.Engine.pack(trajs)
    eng = _engine(model, K)
    fresh = eng.elbo_batch(offsets, X, G, seeds, draws=True, fit=True)
    r_fresh = eng.psis_resample(11, 25, True)
    eng.pin(*pf.Engine.result_arrays(fresh))
    eng.pin(r_fresh["weights"], r_fresh["log_weights"], r_fresh["draws"])
    keep = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in r_fresh.items()}
    elbo0, vh0, draws0 = fresh.elbo.copy(), fresh.fit["vh"].copy(), fresh.draws.copy()
    fresh.elbo[:] = 0; fresh.fit["vh"][:] = 0; fresh.draws[:] = 0
    again = eng.elbo_batch(offsets, X, G, seeds, draws=True, fit=True, into=fresh)
    assert again.elbo is fresh.elbo and again.fit["vh"] is fresh.fit["vh"] and again.draws is fresh.draws
    assert np.array_equal(again.elbo, elbo0) and np.array_equal(again.fit["vh"], vh0, equal_nan=True)
    assert np.array_equal(again.draws, draws0)
    r2 = eng.psis_resample(11, 25, True, into=r_fresh)
    assert r2["weights"] is r_fresh["weights"]
    for k in ("weights", "log_weights", "inds", "ids", "draws"):
        assert np.array_equal(r2[k], keep[k], equal_nan=True)
    assert r2["pareto_k"] == keep["pareto_k"] or (np.isnan(r2["pareto_k"]) and np.isnan(keep["pareto_k"]))
    eng.unpin(*pf.Engine.result_arrays(fresh))
    eng.unpin(r_fresh["weights"], r_fresh["log_weights"], r_fresh["draws"])
    eng.close()


def test_full_size_config3_properties():
    """BASELINE config 3 at its full per-unit size (1024-dim funnel, K = 1000, history 6, real L-BFGS
    trajectories; 6 paths instead of 64 so the test stays in seconds) through size-independent
    properties, plus the oracle on a bounded sample of units."""
    import pathfinder_b200 as pf
    from oracle import pf_oracle as O
    from tests.helpers import make_trajectories

    n, K, J, P = 1024, 1000, 6, 6
    model = pf.Funnel(n)
    trajs = make_trajectories(model, P, seed=20261017, init_scale=10.0, maxiters=25, min_len=8)
    seeds = _seeds(trajs, 4)
    offsets, X, G = pf.Engine.pack(trajs)
    sd = np.concatenate(seeds)
    lean = _engine(model, K, J)
    a = lean.elbo_batch(offsets, X, G, sd, draws=True, per_draw=True, fit=True)
    mm = _engine(model, K, J, materialize_all=True, two_pass=True)
    b = mm.elbo_batch(offsets, X, G, sd, draws=True, per_draw=True, all_draws=True)
    U = a.elbo.size
    # (1) the single-pass statistics formulation == the generic two-pass kernel (algebraic identity)
    fin = np.isfinite(b.elbo)
    assert np.array_equal(fin, np.isfinite(a.elbo))
    scale = np.maximum(1.0, np.nanmax(np.abs(b.logp), axis=0))
    with np.errstate(invalid="ignore"):
        dev = np.nanmax(np.abs(a.logp - b.logp), axis=0) / scale
    assert np.nanmax(dev[fin]) < 1e-7, np.nanmax(dev[fin])   # cancellation in ill-conditioned units, still << 1e-6
    np.testing.assert_allclose(a.logq, b.logq, rtol=1e-12, atol=1e-9)
    # (2) ELBO = mean(logp - logq), SE = sd / sqrt(K)  (src/elbo.jl:16-18)
    logr = b.logp - b.logq
    np.testing.assert_allclose(b.elbo[fin], logr.mean(axis=0)[fin], rtol=1e-10)
    np.testing.assert_allclose(b.elbo_se[fin], (logr.std(axis=0, ddof=1) / np.sqrt(K))[fin], rtol=1e-8)
    # (3) logp of the materialised draws is the model's (host evaluation of the same x)
    for u in (0, U // 2, U - 1):
        ref = O.logp_funnel(b.all_draws[:, :, u])
        ok = np.isfinite(ref)
        np.testing.assert_allclose(b.logp[ok, u], ref[ok], rtol=1e-10, atol=1e-8)
    # (4) K5 regenerates the best iteration's draws bit for bit; argmax follows _findmax_skipnan
    for p in range(P):
        sl = b.unit_slice(p)
        ev = b.elbo[sl]
        assert b.best_iter[p] == O.findmax_skipnan(list(ev))[1]
        ub = sl.start + int(b.best_iter[p]) - 1
        assert np.array_equal(b.draws[:, :, p], b.all_draws[:, :, ub])
        assert np.array_equal(a.draws[:, :, p], b.draws[:, :, p]) or a.best_iter[p] != b.best_iter[p]
    # (5) the oracle on a bounded sample: the first two units of path 0 at full size (they depend on
    # the first three trajectory points only); tolerance policy of _compare_batch: 1e-6, or 50x the
    # oracle's own response to a 1-ulp input perturbation where the history is near-collinear
    from tests.helpers import oracle_batch

    trunc = [(trajs[0][0][:, :3], trajs[0][1][:, :3])]
    sd2 = [seeds[0][:2]]
    orc = oracle_batch(model, trunc, sd2, K, J)
    sens = _sensitivity(model, trunc, sd2, K, J, orc)[0]
    for l in (1, 2):
        est = orc[0]["ests"][l - 1]
        tol = max(RTOL, 50.0 * float(sens[l - 1]))
        assert abs(b.elbo[l - 1] - est["value"]) <= tol * max(1.0, abs(est["value"])), (l, tol)
        assert _rel(b.all_draws[:, :, l - 1], est["draws"]) < tol, (l, tol)
    # (6) pool of P * K draws: weights sum to 1 (test/resample.jl:103-108), resampled columns are
    # pool columns, ids consistent (test/resample.jl:51-59)
    r = lean.psis_resample(9, 500, True)
    assert abs(np.nansum(r["weights"]) - 1.0) < 1e-12 and r["tail_len"] == min(-(-P * K // 5), int(np.ceil(3 * np.sqrt(P * K))))
    pool = a.draws.reshape(n, K * P, order="F")
    assert np.array_equal(r["draws"], pool[:, r["inds"] - 1])
    assert np.array_equal(r["ids"], -(-r["inds"] // K))
    lean.close(); mm.close()


def test_unit_draws_on_demand_equal_the_elbo_stage_payload():
    """pfb_unit_draws regenerates ELBOEstimate.draws / logp / logq (src/elbo.jl:22-29) of any
    iteration: bit-identical to the materialise-all run, for a device family and a GEMM-shaped one."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    n, K = 18, 40
    trajs = [synthetic_trajectory(n, L, 50 + L, scale=0.3) for L in (4, 6)]
    seeds = np.concatenate(_seeds(trajs, 6))
    offsets, X, G = pf.Engine.pack(trajs)
    rng = np.random.default_rng(1)
    Pm = np.linalg.inv(_rand_pd(rng, n))
    for model in (pf.Funnel(n), pf.DenseNormal(rng.normal(size=n), 0.5 * (Pm + Pm.T))):
        full = _engine(model, K, materialize_all=True, two_pass=True)
        a = full.elbo_batch(offsets, X, G, seeds, per_draw=True, all_draws=True)
        lean = _engine(model, K, two_pass=True)
        b = lean.elbo_batch(offsets, X, G, seeds, per_draw=True)
        units = [0, 3, 9, 5]
        d, lp, lq = lean.unit_draws(units)
        assert np.array_equal(d, a.all_draws[:, :, units])
        assert np.array_equal(lq, a.logq[:, units]) and np.array_equal(lp, a.logp[:, units], equal_nan=True)
        assert np.array_equal(lp, b.logp[:, units], equal_nan=True)
        with pytest.raises(pf.PfbError):
            lean.unit_draws([10])
        full.close(); lean.close()


def test_readme_usage_runs():
    """The README's usage snippet at small sizes: device optimiser, resample with / without
    replacement and from fresh draws, a host-closure model, engine-level reuse of pinned outputs."""
    import pathfinder_b200 as pf

    model = pf.Funnel(16)
    res = pf.multipathfinder(model, 100, nruns=6, ndraws_elbo=50, init_scale=3.0, rng=np.random.default_rng(0),
                             optimizer="device", maxiters=30)
    assert res.draws.shape == (16, 100) and len(res.pathfinder_results) == 6
    pr = res.pathfinder_results[0]
    assert pr.draws.shape == (16, 50) and len(pr.elbo_estimates) == len(pr.optim_trace) - 1
    again = pf.resample(res, 60, replace=False)
    assert len(set(again.sample_inds)) == 60
    fresh = pf.resample(res, 60, ndraws_per_run=200)
    assert fresh.pathfinder_results[0].draws.shape == (16, 200)
    host = pf.HostModel(10, logp_batch=lambda X: -0.5 * (X * X).sum(axis=0), grad=lambda x: -x)
    r1 = pf.pathfinder(host, ndraws_elbo=100, ndraws=100, rng=np.random.default_rng(1))
    assert r1.success and np.allclose(r1.fit_distribution.mu, 0.0, atol=1e-6)   # test/singlepath.jl:13-41
    eng = pf.Engine.for_model(model, history_length=6, ndraws_elbo=50)
    offsets, X, G = pf.Engine.pack([(p.optim_trace.points, p.optim_trace.gradients) for p in res.pathfinder_results])
    seeds = np.arange(int(offsets[-1]) - 6, dtype=np.uint64)
    out = eng.elbo_batch(offsets, X, G, seeds, draws=False, fit=True)
    eng.pin(*pf.Engine.result_arrays(out))
    e0 = out.elbo.copy()
    out = eng.elbo_batch(offsets, X, G, seeds, draws=False, fit=True, into=out)
    assert np.array_equal(out.elbo, e0, equal_nan=True)
    d, lp, lq = eng.unit_draws([0, 5])
    assert d.shape == (16, 50, 2)
    eng.unpin(*pf.Engine.result_arrays(out))
    eng.close()


def test_large_n_global_panel_and_ring_mode():
    """n = 2304 (and the odd 2051): K2's panel no longer fits shared memory (global-memory FR rows,
    256-thread CTAs), K1 runs its tall variant, K3 streams the record through the TMA ring — all three
    against the oracle, in both the materialise-all two-pass mode and the lean single-pass mode."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    for n in (2304, 2051):
        model = pf.IsoNormal(n)
        trajs = [synthetic_trajectory(n, L, 5 + L, scale=0.2) for L in (2, 7)]
        _compare_batch(model, trajs, K=40, J=6)
        _lean_vs_oracle(model, trajs, K=40, J=6)


@pytest.mark.parametrize("kind", ["funnel", "dense"])
def test_resampled_columns_regenerated_without_materialising_the_pool(kind):
    """pfb_psis_resample on a batch whose pool draws were never materialised regenerates exactly the
    selected columns (K7r bin + K3 column selection): bit-identical to gathering from the materialised
    pool, with and without replacement, including a path that failed (NaN columns never selected)."""
    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    n, K = 24, 300
    rng = np.random.default_rng(3)
    if kind == "funnel":
        model = pf.Funnel(n)
    else:
        Pm = np.linalg.inv(_rand_pd(rng, n))
        model = pf.DenseNormal(rng.normal(size=n), 0.5 * (Pm + Pm.T))
    trajs = [synthetic_trajectory(n, L, 70 + L, scale=0.3) for L in (3, 0, 6, 4, 9, 2, 5, 7, 8)]
    seeds = np.concatenate(_seeds(trajs, 12))
    offsets, X, G = pf.Engine.pack(trajs)
    ref = _engine(model, K)
    a = ref.elbo_batch(offsets, X, G, seeds, draws=True)              # pool materialised by the download
    lean = _engine(model, K)
    b = lean.elbo_batch(offsets, X, G, seeds, draws=False)            # pool draws never produced
    assert np.array_equal(a.best_iter, b.best_iter)
    pool = a.draws.reshape(n, -1, order="F")
    for replace, nd in ((True, 500), (False, 200)):
        ra = ref.psis_resample(21, nd, True, replace)
        rb = lean.psis_resample(21, nd, True, replace)
        assert np.array_equal(ra["inds"], rb["inds"]) and np.array_equal(ra["weights"], rb["weights"], equal_nan=True)
        assert np.array_equal(rb["draws"], pool[:, rb["inds"] - 1], equal_nan=True)
        assert np.array_equal(ra["draws"], rb["draws"], equal_nan=True)
    # asking for the per-path draws afterwards materialises the pool, identical to the eager one
    d, lp, lq = lean._pool_download()
    assert np.array_equal(d, a.draws, equal_nan=True) and np.array_equal(lp, a.draws_logp, equal_nan=True)
    ref.close(); lean.close()


def test_pool_columns_device_regenerates_only_the_owned_columns():
    """The multi-GPU resample step (pfb_pool_columns_device): global 1-based indices, this engine owns
    [base, base + P K); its columns are regenerated bit-identically, the others are left untouched."""
    import torch

    import pathfinder_b200 as pf
    from tests.helpers import synthetic_trajectory

    n, K = 20, 96
    model = pf.Funnel(n)
    trajs = [synthetic_trajectory(n, L, 90 + L, scale=0.3) for L in (4, 6, 3)]
    seeds = np.concatenate(_seeds(trajs, 13))
    offsets, X, G = pf.Engine.pack(trajs)
    eng = _engine(model, K)
    eng.elbo_batch(offsets, X, G, seeds, draws=False)
    Nloc = 3 * K
    base = 2 * Nloc                                   # "rank 2 of 4"
    rng = np.random.default_rng(0)
    inds = rng.integers(1, 4 * Nloc + 1, size=400).astype(np.int64)
    inds[:5] = [base + 1, base + Nloc, base, base + Nloc + 1, base + 7]   # the edges of the owned range
    d_inds = torch.from_numpy(inds).cuda()
    out = torch.full((400, n), -7.0, dtype=torch.float64, device="cuda")
    eng.pool_columns_device(400, d_inds.data_ptr(), base, out.data_ptr())
    eng.sync()
    got = out.cpu().numpy()
    pool, _, _ = eng._pool_download()                # materialise for comparison
    pool = pool.reshape(n, -1, order="F")
    mine = (inds > base) & (inds <= base + Nloc)
    assert mine[0] and mine[1] and not mine[2] and not mine[3]
    assert np.array_equal(got[mine], pool[:, inds[mine] - 1 - base].T, equal_nan=True)
    assert np.all(got[~mine] == -7.0)
    eng.close()


def test_init_sampler_and_ntasks_keywords():
    """`init_sampler(rng, x)` (src/singlepath.jl:108-110, :332-344) replaces the uniform initialiser,
    also for retries; `ntasks` / `ntasks_per_run` are accepted and change nothing."""
    import pathfinder_b200 as pf

    calls = []

    def sampler(rng, x):
        x[:] = rng.normal(size=x.size) * 0.5
        calls.append(x.copy())

    model = pf.IsoNormal(7)
    a = pf.multipathfinder(model, 40, nruns=3, ndraws_elbo=20, rng=np.random.default_rng(2), init_sampler=sampler,
                           ntasks=4, ntasks_per_run=2)
    assert len(calls) == 3
    for pr, x0 in zip(a.pathfinder_results, calls):
        assert np.array_equal(pr.optim_trace.points[:, 0], x0)
    b = pf.multipathfinder(model, 40, nruns=3, ndraws_elbo=20, rng=np.random.default_rng(2), init_sampler=sampler)
    assert np.array_equal(a.draws, b.draws)
    r = pf.pathfinder(model, ndraws_elbo=10, rng=np.random.default_rng(3), init_sampler=sampler, ntasks=2)
    assert np.array_equal(r.optim_trace.points[:, 0], calls[-1])


def test_gpu_index_streams_are_prefix_consistent():
    """Counter-based RNG on the device: asking for more resampled draws never changes the earlier
    ones (with and without replacement), and the stream depends on the seed."""
    import pathfinder_b200 as pf

    rng = np.random.default_rng(30)
    N, K_run = 4000, 100
    lr = rng.standard_t(4, size=N) * 1.2
    eng = _engine(pf.IsoNormal(4), K_run)
    for replace in (True, False):
        a = eng.psis_resample_host(lr, K_run, 5, 900, True, replace=replace)["inds"]
        b = eng.psis_resample_host(lr, K_run, 5, 150, True, replace=replace)["inds"]
        c = eng.psis_resample_host(lr, K_run, 6, 150, True, replace=replace)["inds"]
        assert np.array_equal(a[:150], b) and not np.array_equal(b, c)
    eng.close()


def test_failed_path_draws_come_from_the_best_fit_with_the_path_seed():
    """src/singlepath.jl:224-228: a failed path (here: every ELBO is -Inf because the target is -Inf on
    half of the space) returns rand(rng, fit_distributions[fit_iteration + 1], ndraws) — fresh draws, not
    the ELBO-stage draws — and they enter the pool with their own log densities."""
    import pathfinder_b200 as pf
    from oracle import pf_oracle as O
    from tests.helpers import synthetic_trajectory

    n, K, J = 14, 64, 6

    def logp_host(X):  # -Inf on a half space that every fitted normal straddles
        with np.errstate(all="ignore"):
            return np.where(np.sum(X, axis=0) > 0.0, -0.5 * np.sum(X * X, axis=0), -np.inf)

    model = pf.HostModel(n, logp_host, lambda x: -x)
    good = pf.IsoNormal(n)
    trajs = [synthetic_trajectory(n, L, 300 + L, scale=0.3) for L in (5, 3)]
    seeds = _seeds(trajs, 12)
    fb = np.array([0xABCDEF0123456789, 77], dtype=np.uint64)
    offsets, X, G = pf.Engine.pack(trajs)
    eng = pf.Engine.for_model(model, J, K, 0)
    eng.set_fallback_seeds(fb)
    res = eng.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=True, per_draw=True, fit=True)
    assert not res.success.any() and np.all(np.isneginf(res.elbo))
    assert np.array_equal(res.best_iter, [1, 1])  # _findmax_skipnan keeps the first of equal values
    for p, (Xp, Gp) in enumerate(trajs):
        mus, Hs, _ = O.fit_mvnormals(Xp, Gp, history_length=J)
        xd, lq = O.failed_path_draws(int(fb[p]), mus[:, 1], Hs[1], K)
        assert _rel(res.draws[:, :, p], xd) < RTOL
        np.testing.assert_allclose(res.draws_logq[:, p], lq, rtol=RTOL, atol=RTOL)
        np.testing.assert_array_equal(res.draws_logp[:, p], logp_host(res.draws[:, :, p]))
        # fresh draws: not the ELBO-stage draws of that iteration
        ue = np.asarray(O.contract_normals(int(seeds[p][0]), n, K))
        xe, _ = O.rand_and_logpdf(ue, mus[:, 1], Hs[1])
        assert not np.allclose(res.draws[:, :, p], xe)
    # the pool: finite weights only where the log ratio is finite; resampled columns are pool columns
    r = eng.psis_resample(5, 40, True)
    logr = (res.draws_logp - res.draws_logq).reshape(-1, order="F")
    assert np.all(r["weights"][np.isneginf(logr)] == 0.0) and abs(r["weights"].sum() - 1.0) < 1e-12
    pool = res.draws.reshape(n, -1, order="F")
    assert np.array_equal(r["draws"], pool[:, r["inds"] - 1])
    eng.close()

    # a registered family on the same inputs succeeds, and its draws ARE the ELBO-stage draws
    eng = pf.Engine.for_model(good, J, K, 0)
    eng.set_fallback_seeds(fb)
    res2 = eng.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=True)
    assert res2.success.all()
    eng.close()


def test_path_without_an_iteration_draws_from_the_identity_fit():
    """L = 0 (the optimiser recorded only the start): fit_iteration = 0 and the reference draws from
    fit_distributions[1] = N(theta_0 + grad_0, I) (src/inverse_hessian.jl:38-40, src/singlepath.jl:224-228)."""
    import pathfinder_b200 as pf
    from oracle import pf_oracle as O
    from tests.helpers import synthetic_trajectory

    n, K, J = 9, 48, 6
    rng = np.random.default_rng(4)
    x0, g0 = rng.normal(size=(n, 1)), rng.normal(size=(n, 1))
    trajs = [synthetic_trajectory(n, 4, 41), (x0, g0)]
    seeds = _seeds(trajs, 2)
    fb = np.array([11, 2**63 + 12345], dtype=np.uint64)
    offsets, X, G = pf.Engine.pack(trajs)
    for model in (pf.Funnel(n), pf.DenseNormal(rng.normal(size=n), np.eye(n) * 2.0)):
        eng = pf.Engine.for_model(model, J, K, 0)
        eng.set_fallback_seeds(fb)
        res = eng.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=True, fit=True)
        assert list(res.success) == [True, False] and res.best_iter[1] == 0
        u = np.asarray(O.contract_normals(int(fb[1]), n, K))
        xd = (x0 + g0) + u
        assert _rel(res.draws[:, :, 1], xd) < 1e-14
        np.testing.assert_allclose(res.draws_logq[:, 1], -(n * np.log(2 * np.pi) + np.sum(u * u, axis=0)) / 2,
                                   rtol=1e-13)
        from tests.helpers import oracle_logp_fn

        np.testing.assert_allclose(res.draws_logp[:, 1], oracle_logp_fn(model)(xd), rtol=1e-10, atol=1e-10)
        # the resample stage regenerates exactly these columns without materialising the pool
        eng2 = pf.Engine.for_model(model, J, K, 0)
        eng2.set_fallback_seeds(fb)
        eng2.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=False)
        r = eng2.psis_resample(3, 30, False)
        pool = res.draws.reshape(n, -1, order="F")
        assert np.array_equal(r["draws"], pool[:, r["inds"] - 1])
        # failures are resolved lazily: run() returns without looking at the success flags, and a resample
        # issued straight behind it (no download in between) runs optimistically, notices the failed path
        # after its own synchronisation and repeats on the corrected pool — same result as above
        eng3 = pf.Engine.for_model(model, J, K, 0)
        eng3.set_fallback_seeds(fb)
        eng3.upload(offsets, X, G, np.concatenate(seeds))
        eng3.run()
        r3 = eng3.psis_resample(3, 30, False)
        assert np.array_equal(r3["inds"], r["inds"]) and np.array_equal(r3["draws"], r["draws"])
        eng3.set_fallback_seeds(fb)
        eng3.run()
        r3w, r2w = eng3.psis_resample(8, 30, True), eng2.psis_resample(8, 30, True)
        assert np.array_equal(r3w["weights"], r2w["weights"]) and np.array_equal(r3w["draws"], r2w["draws"])
        eng.close(); eng2.close(); eng3.close()


def test_nan_log_ratios_get_zero_weight_and_an_all_nan_pool_is_an_error():
    import pathfinder_b200 as pf
    from oracle import psis as OP

    rng = np.random.default_rng(0)
    N = 3000
    logr = rng.normal(size=N) * 2.0
    logr[rng.choice(N, 400, replace=False)] = np.nan
    eng = pf.Engine(4, 0, None, 6, 5, 0)
    r = eng.psis_resample_host(logr, 100, 9, 64, True)
    ref = OP.psis(logr)
    assert np.array_equal(r["weights"], ref["weights"]) and np.array_equal(r["log_weights"], ref["log_weights"])
    assert np.all(r["weights"][np.isnan(logr)] == 0.0) and abs(r["weights"].sum() - 1.0) < 1e-12
    assert np.array_equal(r["inds"], OP.resample_indices(9, ref["weights"], N, 64))
    assert not np.isnan(logr[r["inds"] - 1]).any()
    with pytest.raises(pf.PfbError) as ei:
        eng.psis_resample_host(np.full(200, np.nan), 10, 1, 8, True)
    assert ei.value.code == -5
    eng.close()


def test_unit_fits_export_every_iteration():
    """fit_distributions of ANY iteration on demand (src/singlepath.jl:64 keeps them all)."""
    import pathfinder_b200 as pf
    from oracle import pf_oracle as O
    from tests.helpers import synthetic_trajectory

    n, K, J = 20, 16, 6
    trajs = [synthetic_trajectory(n, L, 70 + L) for L in (7, 3)]
    seeds = _seeds(trajs, 3)
    offsets, X, G = pf.Engine.pack(trajs)
    eng = _engine(pf.IsoNormal(n), K, J)
    res = eng.elbo_batch(offsets, X, G, np.concatenate(seeds), fit=True)
    U = res.elbo.size
    f = eng.unit_fits(np.arange(U))
    for p, (Xp, Gp) in enumerate(trajs):
        mus, Hs, _ = O.fit_mvnormals(Xp, Gp, history_length=J)
        sl = res.unit_slice(p)
        for l in range(1, Xp.shape[1]):
            u = sl.start + l - 1
            W = Hs[l]
            np.testing.assert_allclose(f["mu"][:, u], mus[:, l], rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(f["alpha"][:, u], W.alpha, rtol=1e-10)
            np.testing.assert_allclose(f["logdet"][u], W.logdet(), rtol=1e-9, atol=1e-9)
            assert f["jeff"][u] * 2 == W.k or W.k == n
            k = W.k
            np.testing.assert_allclose(f["vh"][:, :k, u], W.Vh[:, :k], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(f["Vc"][u][:k, :k], W.Vc[:k, :k], rtol=1e-6, atol=1e-9)
        ub = sl.start + int(res.best_iter[p]) - 1
        assert np.array_equal(f["mu"][:, ub], res.fit["mu"][:, p]) and np.array_equal(f["vh"][:, :, ub], res.fit["vh"][:, :, p])
    eng.close()


@pytest.mark.parametrize("J,n", [(13, 80), (16, 20), (24, 130)])
def test_history_longer_than_12_uses_the_generic_kernels(J, n):
    """The reference accepts any history_length (src/inverse_hessian.jl:25).  Beyond 12 the runtime-width
    kernels K2g / K3g take over: same algorithm and conventions, every iteration against the oracle
    (n = 20 < 2J exercises min(n, 2J) reflectors), lean and materialising calls, fit export, resampling."""
    import pathfinder_b200 as pf
    from oracle import psis as OP
    from tests.helpers import synthetic_trajectory

    model = pf.IsoNormal(n)
    trajs = [synthetic_trajectory(n, L, 500 + J + L) for L in (2 * J + 5, 3)]
    res, orc = _compare_batch(model, trajs, K=32, J=J)
    assert max(o["Hs"][-1].k for o in orc) == min(n, 2 * J)
    # funnel target, lean call (no draws): ELBO table vs oracle, then PSIS + regenerated resampled columns
    model = pf.Funnel(n)
    trajs = [synthetic_trajectory(n, 2 * J + 3, 900 + J, scale=0.3), synthetic_trajectory(n, 5, 901 + J, scale=0.3)]
    seeds = _seeds(trajs, 3)
    offsets, X, G = pf.Engine.pack(trajs)
    from tests.helpers import oracle_batch

    orc = oracle_batch(model, trajs, seeds, 40, J)
    lean = _engine(model, 40, J)
    a = lean.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=False, per_draw=True)
    for p, o in enumerate(orc):
        ev = np.array([e["value"] for e in o["ests"]])
        np.testing.assert_allclose(a.elbo[a.unit_slice(p)], ev, rtol=RTOL, atol=RTOL)
    r = lean.psis_resample(4, 30, True)
    full = _engine(model, 40, J)
    b = full.elbo_batch(offsets, X, G, np.concatenate(seeds), draws=True)
    pool = b.draws.reshape(n, -1, order="F")
    assert np.array_equal(r["draws"], pool[:, r["inds"] - 1])
    logr = (b.draws_logp - b.draws_logq).reshape(-1, order="F")
    assert np.array_equal(r["weights"], OP.psis(logr)["weights"])
    lean.close(); full.close()
    # a GEMM-shaped target through K3g + K8g
    rng = np.random.default_rng(J)
    Pm = np.linalg.inv(_rand_pd(rng, n))
    dm = pf.DenseNormal(rng.normal(size=n), 0.5 * (Pm + Pm.T))
    _compare_batch(dm, [synthetic_trajectory(n, 2 * J + 2, 77 + J, scale=0.3)], K=24, J=J)
