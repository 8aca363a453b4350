"""Host-side trajectory producer (stays sequential on the CPU, like src/optimize.jl).

Mirrors ``optimize_with_trace`` (reference: src/optimize.jl:35-59): run L-BFGS on the negative
log density and record, per iteration, the point, the log density and its gradient
(``OptimizationTrace``, src/optimize.jl:110-114), stopping on a non-finite value
(callback rules, src/optimize.jl:103-105).  The reference drives Optim.LBFGS (HagerZhang);
that optimiser is a Julia dependency, so here SciPy's L-BFGS-B with the same memory length
produces the trajectory.  The trajectory is an *input* of the hot path, not part of it.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.optimize import minimize


@dataclass
class OptimizationTrace:
    points: np.ndarray          # [n, L+1]
    log_densities: np.ndarray   # [L+1]
    gradients: np.ndarray       # [n, L+1]  gradients of the LOG density (src/optimize.jl:96)

    def __len__(self):
        return self.points.shape[1]


class _Stop(Exception):
    pass


def optimize_with_trace(model, x0, history_length=6, maxiters=1000, fail_on_nonfinite=True):
    x0 = np.asarray(x0, dtype=np.float64)
    pts, lps, grads = [], [], []

    def record(x):
        with np.errstate(all="ignore"):
            lp = model.logp(x)
            g = model.grad(x)
        # src/optimize.jl:94-105: the callback pushes the point FIRST, then a NaN / +Inf log density
        # or a non-finite gradient ends the run — the offending point stays in the trace
        pts.append(np.array(x, copy=True)); lps.append(lp); grads.append(g)
        bad = np.isnan(lp) or lp == np.inf or not np.all(np.isfinite(g))
        return not (bad and fail_on_nonfinite)

    if not record(x0):
        return OptimizationTrace(np.stack(pts, axis=1), np.array(lps), np.stack(grads, axis=1))

    def fun(x):
        with np.errstate(all="ignore"):
            f, g = -model.logp(x), -model.grad(x)
        if not np.isfinite(f):
            f = 1e300
        return f, np.nan_to_num(g, nan=0.0, posinf=1e300, neginf=-1e300)

    def cb(xk):
        if not record(xk):
            raise _Stop()

    try:
        minimize(fun, x0, jac=True, method="L-BFGS-B", callback=cb,
                 options=dict(maxcor=history_length, maxiter=maxiters, gtol=1e-8, ftol=1e-14, maxls=40))
    except _Stop:
        pass
    return OptimizationTrace(np.stack(pts, axis=1), np.array(lps), np.stack(grads, axis=1))
