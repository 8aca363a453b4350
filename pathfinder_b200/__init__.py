"""pathfinder_b200 — B200-native ELBO-and-resample engine behind the Pathfinder.jl API surface.

Only the hot path of SURVEY.md §8 lives here: `csrc/` (hand-written sm_100a CUDA kernels and
the C ABI, built into libpfb200.so) and the host-side mirror of the reference's
`pathfinder` / `multipathfinder` / `resample` functions that calls it through ctypes.
There is no CPU fallback: without the CUDA library and a GPU every compute call raises.
"""
from ._lib import (PFB_MODEL_DENSENORMAL, PFB_MODEL_DIAGNORMAL, PFB_MODEL_FUNNEL, PFB_MODEL_HLOGISTIC,  # noqa: F401
                   PFB_MODEL_HOSTCALLBACK, PFB_MODEL_ISONORMAL, PfbError)
from .api import (DEFAULT_HISTORY_LENGTH, DEFAULT_NDRAWS_ELBO, ELBOEstimate, FitDistribution,  # noqa: F401
                  MultiPathfinderResult, PathfinderResult, PSISResult, multipathfinder, pathfinder, resample)
from .engine import ElboBatchResult, Engine  # noqa: F401
from .models import DenseNormal, DiagNormal, Funnel, HierLogistic, HostModel, IsoNormal  # noqa: F401
from .optimize import OptimizationTrace, optimize_with_trace  # noqa: F401
