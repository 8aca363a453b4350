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
