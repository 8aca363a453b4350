#!/usr/bin/env python3
"""Run the ELBO stage of a GEMM-shaped target (config 5's dense normal, n = 4096, J = 10, K = 500) on a few
paths — the target of ncu captures of pfb_k8_gemm_logp.

    ncu --set full --import-source on -k regex:pfb_k8_gemm_logp -s 1 -c 1 -o gpurun_out/k8 python scripts/k8_prof.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pathfinder_b200 as pf  # noqa: E402

n, K, J, P = 4096, 500, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 3
model, _ = bench.make_model("dense", n)
trajs, seeds = [], []
for p in range(P):
    rng = np.random.default_rng(bench.MASTER_SEED + p)
    tr = pf.optimize_with_trace(model, (rng.random(n) * 2 - 1) * 2.0, J, 1000)
    trajs.append((tr.points, tr.gradients))
    seeds.append(rng.integers(0, 2**64, size=len(tr) - 1, dtype=np.uint64))
offsets, X, G = pf.Engine.pack(trajs)
eng = pf.Engine(n, model.family, model.blob, J, K, 0)
eng.upload(offsets, X, G, np.concatenate(seeds))
for _ in range(3):
    eng.run()
    eng.sync()
    print(eng.timings(), "units", int(offsets[-1]) - P, flush=True)
