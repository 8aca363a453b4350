"""Registered target-density family (host side).

The reference accepts any Julia closure / LogDensityProblems object (src/singlepath.jl:142-152);
arbitrary host closures cannot run inside a CUDA kernel, so the engine takes a *registered
family id + parameter blob* (include/pfb200.h).  Each class below carries the family id for the
device and a NumPy ``logp`` / ``grad`` used by the host-side L-BFGS (which stays on the CPU,
like src/optimize.jl).
"""
from __future__ import annotations

import numpy as np

from ._lib import PFB_MODEL_DIAGNORMAL, PFB_MODEL_FUNNEL, PFB_MODEL_ISONORMAL


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
