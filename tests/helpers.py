"""Shared helpers for the parity tests: seeded trajectories and the oracle pipeline."""
import numpy as np

from oracle import pf_oracle as O


def make_trajectories(model, P, seed, init_scale, history_length=6, maxiters=200, min_len=0):
    """Host L-BFGS trajectories (inputs of the hot path), seeded."""
    from pathfinder_b200.optimize import optimize_with_trace

    rng = np.random.default_rng(seed)
    out = []
    while len(out) < P:
        x0 = (rng.random(model.n) * 2 - 1) * init_scale
        tr = optimize_with_trace(model, x0, history_length, maxiters)
        if len(tr) - 1 >= min_len:
            out.append((tr.points, tr.gradients))
    return out


def synthetic_trajectory(n, L, seed, scale=1.0):
    """A random smooth 'trajectory' with consistent positive curvature (no optimiser needed)."""
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n)) / np.sqrt(n)
    H = A @ A.T + np.diag(rng.uniform(0.5, 2.0, size=n))  # SPD Hessian of a quadratic -logp
    x = rng.normal(size=n) * scale
    pts, grads = [x.copy()], [-(H @ x)]
    for _ in range(L):
        x = x - rng.uniform(0.05, 0.3) * (H @ x) / np.linalg.norm(H, 2) + 0.01 * rng.normal(size=n)
        pts.append(x.copy())
        grads.append(-(H @ x))
    return np.stack(pts, 1), np.stack(grads, 1)


def oracle_logp_fn(model):
    from pathfinder_b200 import DenseNormal, DiagNormal, Funnel, HierLogistic, HostModel, IsoNormal

    if isinstance(model, HostModel):
        return model.logp_batch

    if isinstance(model, DenseNormal):
        return O.make_logp_dense_gaussian(model.mean, model.prec)
    if isinstance(model, HierLogistic):
        return O.make_logp_hier_logistic(model.X, model.y)

    if isinstance(model, IsoNormal):
        return O.logp_isonormal
    if isinstance(model, Funnel):
        return O.logp_funnel
    if isinstance(model, DiagNormal):
        def f(x):
            z = (x - model.mean[:, None]) / model.sd[:, None]
            return -0.5 * np.sum(z * z, axis=0) - np.sum(np.log(model.sd)) - 0.5 * model.n * np.log(2 * np.pi)
        return f
    raise TypeError(model)


def oracle_batch(model, trajs, seeds_per_path, K, J, normals=None):
    """Run the oracle over a batch; returns per-path dicts."""
    logp_fn = oracle_logp_fn(model)
    res = []
    u0 = 0
    for p, (X, G) in enumerate(trajs):
        mus, Hs, rej = O.fit_mvnormals(X, G, history_length=J)
        L = X.shape[1] - 1
        nrm = None if normals is None else [normals[:, :, u0 + l] for l in range(L)]
        lopt, ests = O.maximize_elbo(seeds_per_path[p], logp_fn, mus, Hs, K, normals=nrm)
        res.append(dict(mus=mus, Hs=Hs, rejected=rej, lopt=lopt, ests=ests,
                        success=O.path_success(L, ests, lopt)))
        u0 += L
    return res
