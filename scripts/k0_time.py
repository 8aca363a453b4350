#!/usr/bin/env python3
"""Time kernel K0 (device L-BFGS) on the bench workload's inits (config 3: 64 paths, 1024-dim funnel)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pathfinder_b200 as pf

n, P = 1024, 64
model = pf.Funnel(n)
x0 = np.stack([(np.random.default_rng(20261017 + p).random(n) * 2 - 1) * 10 for p in range(P)], axis=1)
eng = pf.Engine(n, model.family, None, 6, 1000, 0)
for maxiters in (64, 64, 250, 1000):
    t = time.perf_counter()
    npts, st, nev = eng.lbfgs_batch(x0, maxiters)
    wall = time.perf_counter() - t
    print(f"maxiters {maxiters}: kernel {eng.lbfgs_ms():.3f} ms, call wall {wall*1e3:.2f} ms, points {npts.sum()}, "
          f"evals {nev.sum()}, status {np.bincount(st, minlength=5)}", flush=True)
eng.close()

# GEMM-shaped families at the BASELINE config 4 / 5 shapes
import bench
for name in ("cfg4_hlogistic256_p32_k2000_j6", "cfg5_dense4096_p16_k500_j10"):
    kind, n, P, K, J, scale, nd = bench.CONFIGS[name]
    model, _ = bench.make_model(kind, n)
    x0 = np.stack([(np.random.default_rng(20261017 + p).random(n) * 2 - 1) * scale for p in range(P)], axis=1)
    eng = pf.Engine.for_model(model, J, 8, 0)
    for rep in range(2):
        npts, st, nev = eng.lbfgs_batch(x0, 64)
    print(f"{name}: K0 {eng.lbfgs_ms():.2f} ms, points {npts.sum()}, evals {nev.sum()}, status {np.bincount(st, minlength=5)}",
          flush=True)
    eng.close()
