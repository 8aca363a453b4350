"""Host-side mirror of the reference's public API for the accelerated path.

Same names, keyword meaning and failure behaviour as Pathfinder.jl (paths relative to the
reference repo):

  pathfinder        src/singlepath.jl:101-257     PathfinderResult       src/singlepath.jl:53-70
  multipathfinder   src/multipath.jl:94-245       MultiPathfinderResult  src/multipath.jl:31-44
  resample          src/resample.jl:20-46

The sequential L-BFGS stays on the host (optimize.py); everything the reference does after it
(`fit_mvnormals`, `maximize_elbo`, draw selection, `_compute_psis_result`, `_resample`) is one
batched call into libpfb200.so.  The per-path calls of `_chunk_tmap` (src/multipath.jl:190-208)
are hoisted: all trajectories first, one engine call, then result assembly; paths that fail
(src/singlepath.jl:309-314) are re-initialised and sent as a further batch, up to `ntries`
(src/singlepath.jl:259-283).
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field

import numpy as np

from . import distributed as D
from .engine import Engine
from .optimize import OptimizationTrace, optimize_with_trace

DEFAULT_HISTORY_LENGTH = 6  # src/Pathfinder.jl:24
DEFAULT_NDRAWS_ELBO = 5     # src/Pathfinder.jl:27


@dataclass
class ELBOEstimate:
    """src/elbo.jl:22-29 (draws / per-draw densities are kept for the best iteration only; while the
    batch is current, `Engine.unit_draws` regenerates any other iteration's bit for bit from its
    seed — pfb_unit_draws)."""

    value: float
    std_err: float


@dataclass
class FitDistribution:
    """MvNormal(mu, Sigma) with Sigma = WoodburyPDMat in factored form (src/woodbury.jl:259)."""

    mu: np.ndarray
    alpha: np.ndarray     # diag(A)
    vh: np.ndarray        # n x KP Householder reflectors of F.Q
    T: np.ndarray         # KP x KP compact-WY factor of F.Q
    Vc: np.ndarray        # KP x KP upper Cholesky factor F.V
    logdet: float
    history_length_effective: int


class PathfinderResult:
    """src/singlepath.jl:53-70.  `draws` (and their log densities) may still be on the device when
    the result is built (`multipathfinder(..., engine=...)` on a caller-owned engine): they are
    fetched on first access, or when the engine's pool is about to be overwritten."""

    def __init__(self, input, rng, fit_distribution, draws, fit_iteration, num_tries, optim_trace, elbo_estimates,
                 num_bfgs_updates_rejected, success=True, draws_logp=None, draws_logq=None, lazy=None):
        self.input = input
        self.rng = rng
        self.fit_distribution = fit_distribution
        self._draws = draws                  # [n, ndraws]
        self.fit_iteration = fit_iteration   # 1-based; 0 = failed before any iteration
        self.num_tries = num_tries
        self.optim_trace = optim_trace
        self.elbo_estimates = elbo_estimates
        self.num_bfgs_updates_rejected = num_bfgs_updates_rejected
        self.success = success
        self._draws_logp = draws_logp
        self._draws_logq = draws_logq
        self._lazy = lazy                    # (ElboBatchResult, path index, ndraws)

    def _fetch(self):
        if self._lazy is not None:
            res, j, nd = self._lazy
            self._lazy = None
            res.fetch_draws()
            self._draws = res.draws[:, :nd, j]
            self._draws_logp = res.draws_logp[:nd, j]
            self._draws_logq = res.draws_logq[:nd, j]

    @property
    def draws(self):
        self._fetch()
        return self._draws

    @property
    def draws_logp(self):
        self._fetch()
        return self._draws_logp

    @property
    def draws_logq(self):
        self._fetch()
        return self._draws_logq


@dataclass
class PSISResult:
    log_weights: np.ndarray
    weights: np.ndarray
    pareto_shape: float
    tail_length: int


@dataclass
class MultiPathfinderResult:
    input: object
    rng: object
    draws: np.ndarray                  # [n, ndraws]
    draw_component_ids: np.ndarray     # [ndraws] 1-based
    pathfinder_results: list
    psis_result: PSISResult | None
    sample_inds: np.ndarray = field(default=None, repr=False)
    engine: Engine = field(default=None, repr=False)


def _uniform_init(rng, n, scale, init_sampler=None):
    """UniformSampler (src/singlepath.jl:332-344): iid U[-scale, scale]; or the caller's
    `init_sampler(rng, x)` filling `x` in place like the reference's keyword (src/singlepath.jl:108-110)."""
    if init_sampler is not None:
        x = np.empty(n, dtype=np.float64)
        out = init_sampler(rng, x)
        return np.asarray(x if out is None else out, dtype=np.float64)
    return (rng.random(n) * 2.0 - 1.0) * scale


def _draw_seeds(rng, m):
    return rng.integers(0, 2**64, size=m, dtype=np.uint64)


DEVICE_LBFGS_FAMILIES = (0, 1, 2, 3, 4)  # every registered device-side family (include/pfb200.h)


def _use_device_optimizer(model, optimizer):
    if optimizer not in ("auto", "device", "host"):
        raise ValueError("optimizer must be 'auto', 'device' or 'host'")
    if optimizer == "device" and model.family not in DEVICE_LBFGS_FAMILIES:
        raise ValueError("the device L-BFGS does not cover this family; use optimizer='host'")
    return optimizer == "device" or (optimizer == "auto" and model.family in DEVICE_LBFGS_FAMILIES)


def _run_paths(engine, model, inits, path_rngs, *, history_length, maxiters, ntries, init_scale, ndraws_run=None,
               optimizer="host", gtol=1e-8, ftol=1e-14, lazy_draws=False, init_sampler=None):
    """Optimise every path (host L-BFGS, or kernel K0 for the closed-form families), run the ELBO
    stage as one batch, retry failures.  lazy_draws: leave the best-iteration draws on the device."""
    want_draws = "lazy" if (lazy_draws and not (ndraws_run is not None and ndraws_run > engine.K)) else True
    device_opt = _use_device_optimizer(model, optimizer)
    P = len(inits)
    if P == 0:  # a rank that owns no run (nruns < world size): nothing to optimise, an empty pool
        return [], None
    final = [None] * P
    todo = list(range(P))
    tries = [0] * P
    cur_init = list(inits)
    last = None
    while todo:
        traces, seeds = [], []
        if device_opt:
            # K0: all trajectories in one launch; they stay on the device for K1/K2
            for p in todo:
                tries[p] += 1
            x0s = np.stack([np.asarray(cur_init[p], dtype=np.float64) for p in todo], axis=1)
            npts, _, _ = engine.lbfgs_batch(x0s, maxiters, None, gtol, ftol)
            seeds = [_draw_seeds(path_rngs[p], int(npts[j]) - 1) for j, p in enumerate(todo)]  # src/elbo.jl:2
            # a failed path's draws are rand(rng, fit_distribution, ndraws) with the PATH's rng (src/singlepath.jl:226-228)
            engine.set_fallback_seeds([int(_draw_seeds(path_rngs[p], 1)[0]) for p in todo])
            engine.batch_from_lbfgs(np.concatenate(seeds) if seeds else np.zeros(0, np.uint64))
            engine.run()
            res = engine.download(draws=want_draws, fit=True)
            off, Xd, FXd, Gd = engine.lbfgs_download()
            traces = [OptimizationTrace(Xd[:, off[j]:off[j + 1]], FXd[off[j]:off[j + 1]], Gd[:, off[j]:off[j + 1]])
                      for j in range(len(todo))]
        else:
            for p in todo:
                tries[p] += 1
                tr = optimize_with_trace(model, cur_init[p], history_length, maxiters)
                traces.append(tr)  # a non-finite start leaves a 1-point trace (L = 0): the path fails
                seeds.append(_draw_seeds(path_rngs[p], len(tr) - 1))  # src/elbo.jl:2
            offsets, X, G = Engine.pack([(t.points, t.gradients) for t in traces])
            engine.set_fallback_seeds([int(_draw_seeds(path_rngs[p], 1)[0]) for p in todo])  # src/singlepath.jl:226-228
            res = engine.elbo_batch(offsets, X, G, np.concatenate(seeds) if seeds else np.zeros(0, np.uint64),
                                    draws=want_draws, fit=True)
        if ndraws_run is not None and ndraws_run > engine.K:
            # top-up draws from the fitted normal with the path's rng (src/singlepath.jl:228-230)
            top_seeds = np.array([int(_draw_seeds(path_rngs[p], 1)[0]) for p in todo], dtype=np.uint64)
            xd, lp, lq = engine.draw_from_fits(ndraws_run - engine.K, top_seeds)
            res.draws = np.asfortranarray(np.concatenate([res.draws, xd], axis=1))
            res.draws_logp = np.asfortranarray(np.concatenate([res.draws_logp, lp], axis=0))
            res.draws_logq = np.asfortranarray(np.concatenate([res.draws_logq, lq], axis=0))
            res.topped_up = True
        last = (todo[:], res, traces)
        retry = []
        for j, p in enumerate(todo):
            ok = bool(res.success[j])
            if ok or tries[p] >= ntries:
                final[p] = (j, res, traces[j], tries[p])
            else:
                cur_init[p] = _uniform_init(path_rngs[p], model.n, init_scale, init_sampler)  # src/singlepath.jl:278
                retry.append(p)
        todo = retry
    return final, last


def _assemble_path(model, rng, entry, ndraws, K):
    j, res, trace, ntry = entry
    sl = res.unit_slice(j)
    ests = [ELBOEstimate(float(v), float(s)) for v, s in zip(res.elbo[sl], res.elbo_se[sl])]
    ok = bool(res.success[j])
    if not ok:
        warnings.warn(f"Pathfinder failed after {ntry} tries. Increase `ntries`, inspect the model for "
                      "numerical instability, or provide a more suitable `init_sampler`.")
    rej = int(res.n_rejected[j])
    if rej > 0:
        perc = round(rej * 100.0 / len(trace), 1)
        warnings.warn(f"{rej} ({perc}%) updates to the inverse Hessian estimate were rejected to keep it "
                      "positive definite.")
    fit = None
    if res.fit is not None and res.best_iter[j] > 0:
        f = res.fit
        fit = FitDistribution(f["mu"][:, j].copy(), f["alpha"][:, j].copy(), f["vh"][:, :, j].copy(),
                              f["T"][j].copy(), f["Vc"][j].copy(), float(f["logdet"][j]), int(f["jeff"][j]))
    elif res.fit is not None and len(trace) >= 1:
        # no iteration at all: fit_distributions[1] = N(theta_0 + H_0 grad_0, H_0), H_0 = I
        # (src/inverse_hessian.jl:38-40, src/mvnormal.jl:17) — the normal a failed path's draws come from
        KP = res.fit["vh"].shape[1]
        n = trace.points.shape[0]
        fit = FitDistribution(trace.points[:, 0] + trace.gradients[:, 0], np.ones(n), np.zeros((n, KP)),
                              np.zeros((KP, KP)), np.zeros((KP, KP)), 0.0, 0)
    if res.draws is None:  # still on the device (lazy): fetched on first access
        return PathfinderResult(model, rng, fit, None, int(res.best_iter[j]), ntry, trace, ests, rej, ok,
                                lazy=(res, j, ndraws))
    # views into the batch's download buffers (F-order, so each path's slab is contiguous): no copies
    draws = res.draws[:, :ndraws, j]
    return PathfinderResult(model, rng, fit, draws, int(res.best_iter[j]), ntry, trace, ests, rej, ok,
                            res.draws_logp[:ndraws, j], res.draws_logq[:ndraws, j])


def pathfinder(model, *, init=None, init_scale=2.0, init_sampler=None, ndraws_elbo=DEFAULT_NDRAWS_ELBO, ndraws=None,
               rng=None, history_length=DEFAULT_HISTORY_LENGTH, ntries=1000, maxiters=1000, device=0, engine=None,
               optimizer="host", ntasks=None):
    """Single-path Pathfinder (src/singlepath.jl:101-139).  optimizer: 'host' (SciPy L-BFGS on the
    CPU, like src/optimize.jl), 'device' (kernel K0, registered families) or 'auto'.  `ntasks` is
    accepted for signature compatibility: results never depend on it (src/singlepath.jl:114-117), the
    device batches over iterations instead of tasks."""
    rng = np.random.default_rng() if rng is None else rng
    ndraws = ndraws_elbo if ndraws is None else ndraws
    x0 = _uniform_init(rng, model.n, init_scale, init_sampler) if init is None else np.asarray(init, dtype=np.float64)
    if x0.shape != (model.n,):
        raise ValueError("init has the wrong dimension")
    own = engine is None
    if own:
        engine = Engine.for_model(model, history_length, ndraws_elbo, device)
    try:
        final, _ = _run_paths(engine, model, [x0], [rng], history_length=history_length, maxiters=maxiters,
                              ntries=ntries, init_scale=init_scale, ndraws_run=ndraws, optimizer=optimizer,
                              init_sampler=init_sampler)
        return _assemble_path(model, rng, final[0], ndraws, ndraws_elbo)
    finally:
        if own:
            engine.close()


def multipathfinder(model, ndraws, *, nruns=None, init=None, ndraws_elbo=DEFAULT_NDRAWS_ELBO,
                    ndraws_per_run=None, importance=True, rng=None, history_length=DEFAULT_HISTORY_LENGTH,
                    init_scale=2.0, init_sampler=None, ntries=1000, maxiters=1000, device=0, engine=None, group=None,
                    optimizer="host", ntasks=None, ntasks_per_run=None):
    """Multi-path Pathfinder (src/multipath.jl:94-245).  `ntasks` / `ntasks_per_run` are accepted for
    signature compatibility (results never depend on them, src/multipath.jl:104-108).

    Under an initialised ``torch.distributed`` process group (one process per GPU) the runs shard
    across the ranks (distributed.py): every rank must call with the same arguments and an
    identically seeded ``rng``; every rank returns the same draws / ids / PSIS result, and
    ``pathfinder_results`` holds the rank's own runs."""
    rng = np.random.default_rng() if rng is None else rng
    if init is None:
        if nruns is None or nruns <= 0:
            raise ValueError("A positive `nruns` must be set or `init` must be provided.")  # :146-148
        inits = [None] * nruns
    else:
        inits = [np.asarray(x, dtype=np.float64) for x in init]
    nruns = len(inits)
    if ndraws_per_run is None:
        ndraws_per_run = max(ndraws_elbo, -(-ndraws // max(nruns, 1)))  # src/multipath.jl:138
    if ndraws > ndraws_per_run * nruns:
        warnings.warn("More draws requested than total number of draws across replicas. Draws will not be unique.")
    run_seeds = _draw_seeds(rng, nruns)  # src/multipath.jl:162
    path_rngs = [np.random.Generator(np.random.Philox(key=int(s))) for s in run_seeds]
    inits = [(_uniform_init(path_rngs[p], model.n, init_scale, init_sampler) if x is None else x)
             for p, x in enumerate(inits)]
    seed = int(_draw_seeds(rng, 1)[0])
    dist_on = group is not False and D.is_distributed(group)  # group=False: this process alone, whatever is initialised
    lo, hi = 0, nruns
    if dist_on:
        import torch.distributed as dist

        lo, hi = D.shard_range(nruns, dist.get_rank(group), dist.get_world_size(group))
        import torch

        if device == 0 and dist.get_backend(group) == "nccl":
            device = torch.cuda.current_device()  # one process per GPU: the rank's device
    own = engine is None
    if own:
        engine = Engine.for_model(model, history_length, ndraws_elbo, device)
    # on a caller-owned engine the per-path draws stay device-resident until looked at (the engine
    # hands them over before its pool is overwritten or freed); an engine owned by this call is
    # closed on return, so its draws come back eagerly
    final, last = _run_paths(engine, model, inits[lo:hi], path_rngs[lo:hi], history_length=history_length,
                             maxiters=maxiters, ntries=ntries, init_scale=init_scale, ndraws_run=ndraws_per_run,
                             optimizer=optimizer, lazy_draws=not own, init_sampler=init_sampler)
    results = [_assemble_path(model, path_rngs[lo + j], final[j], ndraws_per_run, ndraws_elbo)
               for j in range(hi - lo)]
    # PSIS pool: draw-fastest, component-slowest (test/resample.jl:81-88)
    K_run = ndraws_per_run
    if dist_on:
        # the exchange lives behind the C ABI (pfb_pool_exchange_resample): all-gather of the pools' log
        # densities, PSIS + index draw replicated, owned columns regenerated on their rank, sum-reduce.
        # Nothing of the pool crosses PCIe unless it had to be assembled on the host (retries, top-up).
        counts = D.shard_counts(nruns, dist.get_world_size(group))
        if getattr(engine, "_comm", None) != (dist.get_rank(group), dist.get_world_size(group)):
            engine.comm_init(group)
        P_loc = hi - lo
        single_batch = (last is not None and len(last[0]) == P_loc and K_run == ndraws_elbo
                        and not getattr(last[1], "topped_up", False))
        if not single_batch:
            if results:
                engine.pool_set(P_loc, K_run, np.stack([pr.draws for pr in results], axis=2),
                                np.stack([pr.draws_logp for pr in results], axis=1),
                                np.stack([pr.draws_logq for pr in results], axis=1))
            else:
                engine.pool_set(0, K_run, None, np.zeros((K_run, 0)), np.zeros((K_run, 0)))
        r = engine.pool_exchange_resample(counts, seed, ndraws, importance)
    else:
        single_batch = (last is not None and len(last[0]) == nruns and K_run == ndraws_elbo
                        and not getattr(last[1], "topped_up", False))
        if single_batch:
            r = engine.psis_resample(seed, ndraws, importance)  # pool still resident on the device
        else:
            pool = np.concatenate([pr.draws for pr in results], axis=1)
            logr = np.concatenate([pr.draws_logp - pr.draws_logq for pr in results])
            r = engine.psis_resample_host(logr if importance else None, K_run, seed, ndraws, importance, pool=pool)
    psis = PSISResult(r["log_weights"], r["weights"], r["pareto_k"], r["tail_len"]) if importance else None
    return MultiPathfinderResult(model, rng, r["draws"], r["ids"], results, psis, r["inds"],
                                 engine if not own else _close_and_none(engine))


def _close_and_none(engine):
    engine.close()
    return None


def resample(result: MultiPathfinderResult, ndraws, *, rng=None, replace=True, importance=True,
             ndraws_per_run=None, device=0, history_length=DEFAULT_HISTORY_LENGTH):
    """Re-resample a fitted result (src/resample.jl:20-46): from its stored draws, or — with
    `ndraws_per_run` — from fresh draws of every path's fitted normal (src/resample.jl:102-109),
    rebuilt on the device from the stored trajectories without an ELBO stage."""
    rng = result.rng if rng is None else rng
    prs = result.pathfinder_results
    model = result.input
    if ndraws_per_run is None:
        pool = np.concatenate([pr.draws for pr in prs], axis=1)
        K_run = prs[0].draws.shape[1]
        seed = int(_draw_seeds(rng, 1)[0])
        eng = Engine.for_model(model, history_length, K_run, device)
        try:
            if importance:
                # the log ratios of the stored draws are the ELBO stage's logp - logq (what
                # _compute_log_importance_ratios, src/resample.jl:81-95, recomputes)
                logr = np.concatenate([pr.draws_logp - pr.draws_logq for pr in prs])
                r = eng.psis_resample_host(logr, K_run, seed, ndraws, True, pool=pool, replace=replace)
                psis = PSISResult(r["log_weights"], r["weights"], r["pareto_k"], r["tail_len"])
            else:
                r = eng.psis_resample_host(None, K_run, seed, ndraws, False, pool=pool, replace=replace)
                psis = None
        finally:
            eng.close()
        return MultiPathfinderResult(model, rng, r["draws"], r["ids"], prs, psis, r["inds"], None)
    K_run = int(ndraws_per_run)
    eng = Engine.for_model(model, history_length, K_run, device)
    try:
        offsets, X, G = Engine.pack([(pr.optim_trace.points, pr.optim_trace.gradients) for pr in prs])
        U = int(offsets[-1]) - len(prs)
        eng.upload(offsets, X, G, np.zeros(U, dtype=np.uint64))
        eng.fit_only([pr.fit_iteration for pr in prs])
        seeds = _draw_seeds(rng, len(prs))
        xd, lp, lq = eng.draw_from_fits(K_run, seeds, keep_as_pool=True)
        r = eng.psis_resample(int(_draw_seeds(rng, 1)[0]), ndraws, importance, replace)
    finally:
        eng.close()
    new_prs = []
    for j, pr in enumerate(prs):
        q = PathfinderResult(pr.input, pr.rng, pr.fit_distribution, xd[:, :, j].copy(), pr.fit_iteration,
                             pr.num_tries, pr.optim_trace, pr.elbo_estimates, pr.num_bfgs_updates_rejected,
                             pr.success, lp[:, j].copy(), lq[:, j].copy())
        new_prs.append(q)
    psis = PSISResult(r["log_weights"], r["weights"], r["pareto_k"], r["tail_len"]) if importance else None
    return MultiPathfinderResult(model, rng, r["draws"], r["ids"], new_prs, psis, r["inds"], None)
