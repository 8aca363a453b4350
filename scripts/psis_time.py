#!/usr/bin/env python3
"""Time the PSIS + resample stage (K6 + K7) on device-resident log densities for pool sizes N."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pathfinder_b200 as pf
from pathfinder_b200 import _lib
from pathfinder_b200._lib import pfb_resample_out

eng = pf.Engine(8, 0, None, 6, 1000, 0)
lib = eng.lib
for N in (64000, 128000, 256000, 512000):
    g = torch.Generator(device="cuda").manual_seed(1)
    lp = torch.randn(N, dtype=torch.float64, device="cuda", generator=g) * 3
    lq = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    inds = np.empty(1000, dtype=np.int64); ids = np.empty(1000, dtype=np.int64)
    def call(full):
        out = pfb_resample_out()
        out.inds = inds.ctypes.data_as(C.c_void_p); out.ids = ids.ctypes.data_as(C.c_void_p)
        if full:
            w = np.empty(N); lw = np.empty(N)
            out.weights = w.ctypes.data_as(C.c_void_p); out.log_weights = lw.ctypes.data_as(C.c_void_p)
        rc = lib.pfb_psis_resample_device(eng.h, 8, N, 1000, C.c_void_p(lp.data_ptr()), C.c_void_p(lq.data_ptr()), None,
                                          C.c_uint64(5), 1000, 1, 1, C.byref(out))
        assert rc == 0
    for full in (False, True):
        call(full); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5): call(full)
        torch.cuda.synchronize()
        print(f"N={N} weights_to_host={full}: {(time.perf_counter()-t)/5*1e3:.3f} ms per call", flush=True)
eng.close()
