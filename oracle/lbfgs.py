"""oracle/lbfgs.py — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU side of the engine's L-BFGS trajectory contract (pathfinder_b200/csrc/pf_lbfgs.h; the
trajectory producer of src/optimize.jl:35-59, default_optimizer src/Pathfinder.jl:29-35).  The
algorithm is the shared header compiled by g++ (oracle/pforacle_lbfgs.cpp) with a host context
that emulates the device's reduction order, so kernel K0 must reproduce these trajectories bit
for bit.  Parity with Optim.jl's own iterates is unpinned (third-party Julia code, absent from
/root/reference): what this oracle pins is CPU == GPU, and `tests/test_oracle_cpu.py` checks the
optimiser itself against the analytic optima and SciPy's L-BFGS-B.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

FAMILY_ISONORMAL, FAMILY_FUNNEL, FAMILY_DIAGNORMAL, FAMILY_DENSENORMAL, FAMILY_HLOGISTIC = 0, 1, 2, 3, 4
STATUS = {0: "gtol", 1: "ftol", 2: "maxiters", 3: "linesearch", 4: "nonfinite"}


def clib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_build", "libpforacle_lbfgs.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = ctypes.CDLL(path)
        lib.pfo_lbfgs_path.restype = ctypes.c_int
        lib.pfo_lbfgs_path.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


def hlogistic_c0(n):
    """Constant of the hierarchical-logistic prior: -(n / 2) log(2 pi) - log(2.5), as pfb_register_model
    forms it."""
    return -0.5 * n * 1.8378770664093453 - float(np.log(2.5))


def lbfgs_path(family, x0, history_length=6, maxiters=1000, max_points=None, gtol=1e-8, ftol=1e-14,
               mean=None, sd=None, prec=None, Xobs=None, yobs=None):
    """One trajectory: returns (points [n, L+1], log_densities [L+1], gradients [n, L+1], status, nevals)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.size
    max_points = min(maxiters + 1, max_points or maxiters + 1)
    X = np.zeros((n, max_points), order="F")
    G = np.zeros((n, max_points), order="F")
    FX = np.zeros(max_points)
    st, nev = ctypes.c_int(0), ctypes.c_int(0)
    mp0 = mp1 = None
    c0 = 0.0
    if family == FAMILY_DIAGNORMAL:
        mp0 = np.ascontiguousarray(mean, dtype=np.float64)
        sd = np.ascontiguousarray(sd, dtype=np.float64)
        mp1 = 1.0 / sd
        # the same summation order as pfb_register_model (pfb_api.cu)
        c0 = -0.5 * n * 1.8378770664093453
        for v in sd:
            c0 -= np.log(v)
    if family == FAMILY_DENSENORMAL:
        mp0 = np.ascontiguousarray(mean, dtype=np.float64)
        mp1 = np.asfortranarray(prec, dtype=np.float64)
    nobs = 0
    mp2 = None
    if family == FAMILY_HLOGISTIC:
        mp0 = np.asfortranarray(Xobs, dtype=np.float64)
        mp1 = np.ascontiguousarray(yobs, dtype=np.float64)
        mp2 = np.ascontiguousarray(Xobs, dtype=np.float64)  # row-major X == column-major X'
        nobs = mp0.shape[0]
        c0 = hlogistic_c0(n)
    np_ = clib().pfo_lbfgs_path(int(family), n, int(nobs), None if mp0 is None else mp0.ctypes.data,
                                None if mp1 is None else mp1.ctypes.data, None if mp2 is None else mp2.ctypes.data,
                                float(c0), int(history_length),
                                int(maxiters), int(max_points), float(gtol), float(ftol), x0.ctypes.data,
                                X.ctypes.data, G.ctypes.data, FX.ctypes.data, ctypes.byref(st), ctypes.byref(nev))
    return (np.asfortranarray(X[:, :np_]), FX[:np_].copy(), np.asfortranarray(G[:, :np_]), st.value, nev.value)
