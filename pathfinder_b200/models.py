"""Registered target-density family (host side).

The reference accepts any Julia closure / LogDensityProblems object (src/singlepath.jl:142-152);
arbitrary host closures cannot run inside a CUDA kernel, so the engine takes a *registered
family id + parameter blob* (include/pfb200.h).  Each class below carries the family id for the
device and a NumPy ``logp`` / ``grad`` used by the host-side L-BFGS (which stays on the CPU,
like src/optimize.jl).
"""
from __future__ import annotations

import numpy as np

from ._lib import (PFB_MODEL_DENSENORMAL, PFB_MODEL_DIAGNORMAL, PFB_MODEL_FUNNEL, PFB_MODEL_HLOGISTIC,
                   PFB_MODEL_HOSTCALLBACK, PFB_MODEL_ISONORMAL)


class IsoNormal:
    """logp(x) = -sum(abs2, x) / 2          (test/singlepath.jl:15)"""

    family = PFB_MODEL_ISONORMAL

    def __init__(self, n):
        self.n = int(n)
        self.blob = None

    def logp(self, x):
        return -0.5 * float(np.dot(x, x))

    def grad(self, x):
        return -np.asarray(x, dtype=np.float64)


class Funnel:
    """Neal's funnel exactly as docs/src/examples/quickstart.md:229-234."""

    family = PFB_MODEL_FUNNEL

    def __init__(self, n):
        self.n = int(n)
        self.blob = None

    def logp(self, x):
        n = self.n
        tau = x[0]
        ss = float(np.dot(x[1:], x[1:]))
        return ((tau / 3.0) ** 2 + (n - 1) * tau + np.exp(-tau) * ss) / -2.0

    def grad(self, x):
        n = self.n
        tau = x[0]
        e = np.exp(-tau)
        g = np.empty(n)
        g[0] = -(2.0 * tau / 9.0 + (n - 1) - e * float(np.dot(x[1:], x[1:]))) / 2.0
        g[1:] = -e * x[1:]
        return g


class DiagNormal:
    """Independent normals N(mean_i, sd_i^2) (normalised), e.g. the 1-D target of test/elbo.jl:8-12."""

    family = PFB_MODEL_DIAGNORMAL

    def __init__(self, mean, sd):
        self.mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
        self.sd = np.atleast_1d(np.asarray(sd, dtype=np.float64))
        self.n = self.mean.size
        self.blob = np.concatenate([self.mean, self.sd])
        self._c0 = -np.sum(np.log(self.sd)) - 0.5 * self.n * np.log(2 * np.pi)

    def logp(self, x):
        z = (x - self.mean) / self.sd
        return -0.5 * float(np.dot(z, z)) + self._c0

    def grad(self, x):
        return -(x - self.mean) / self.sd**2


class DenseNormal:
    """logp(x) = -(x - m)' P (x - m) / 2   (docs/src/examples/quickstart.md:17-24; BASELINE config 5).
    The device evaluates it as a GEMM over the materialised draws (K8)."""

    family = PFB_MODEL_DENSENORMAL

    def __init__(self, mean, prec):
        self.mean = np.asarray(mean, dtype=np.float64)
        self.prec = np.asarray(prec, dtype=np.float64)
        self.n = self.mean.size
        if self.prec.shape != (self.n, self.n):
            raise ValueError("prec must be n x n")
        self.blob = np.concatenate([self.mean, np.asfortranarray(self.prec).reshape(-1, order="F")])

    def logp(self, x):
        z = x - self.mean
        return -0.5 * float(z @ (self.prec @ z))

    def grad(self, x):
        return -(self.prec @ (x - self.mean))


class HierLogistic:
    """Hierarchical logistic regression: theta = (log tau, b0, b_1..b_p),
    log tau ~ N(0, 1), b0 ~ N(0, 2.5^2), b_j ~ N(0, tau^2), y_i ~ Bernoulli(sigmoid(b0 + x_i'b))."""

    family = PFB_MODEL_HLOGISTIC

    def __init__(self, X, y):
        self.X = np.asarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.nobs, p = self.X.shape
        self.n = p + 2
        self.blob = np.concatenate([[float(self.nobs)], np.asfortranarray(self.X).reshape(-1, order="F"), self.y])

    def logp(self, th):
        lt, b0, b = th[0], th[1], th[2:]
        eta = b0 + self.X @ b
        ll = float(np.sum(self.y * eta - np.logaddexp(0.0, eta)))
        h = 0.5 * np.log(2 * np.pi)
        lp = -0.5 * lt * lt - h
        lp += -0.5 * (b0 / 2.5) ** 2 - np.log(2.5) - h
        lp += -0.5 * float(b @ b) * np.exp(-2 * lt) - b.size * lt - b.size * h
        return lp + ll

    def grad(self, th):
        lt, b0, b = th[0], th[1], th[2:]
        eta = b0 + self.X @ b
        r = self.y - 1.0 / (1.0 + np.exp(-eta))
        g = np.empty_like(th)
        e2 = np.exp(-2 * lt)
        g[0] = -lt + float(b @ b) * e2 - b.size
        g[1] = -b0 / 6.25 + float(np.sum(r))
        g[2:] = -b * e2 + self.X.T @ r
        return g


class HostModel:
    """Arbitrary target density evaluated on the HOST: what the reference accepts
    everywhere — a closure / LogDensityProblems object (src/singlepath.jl:142-152, :186).

    ``logp_batch(X)`` maps an ``n x m`` array of draws (columns) to ``m`` log densities — the batched
    form of ``logp.(eachcol(x))`` (src/elbo.jl:15); ``grad(x)`` is the gradient of the log density
    for the host L-BFGS.  The device streams draw tiles to pinned memory and overlaps the callback
    with sampling the next tile."""

    family = PFB_MODEL_HOSTCALLBACK
    blob = None

    def __init__(self, n, logp_batch, grad, logp=None):
        self.n = int(n)
        self.logp_batch = logp_batch
        self._grad = grad
        self._logp = logp

    def logp(self, x):
        if self._logp is not None:
            return float(self._logp(x))
        return float(np.asarray(self.logp_batch(np.asarray(x, dtype=np.float64)[:, None])).reshape(-1)[0])

    def grad(self, x):
        return np.asarray(self._grad(x), dtype=np.float64)
