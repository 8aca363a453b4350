#!/usr/bin/env python3
"""Run the ELBO stage of the bench workload (config 3) a few times — the target of ncu captures.

    ncu --set full --import-source on -k regex:pfb_k3_elbo_sample -s 3 -c 1 -o gpurun_out/x python scripts/k3_prof.py [lean|two_pass|m]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pathfinder_b200 as pf  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "lean"
cache = "/tmp/k3_ab_workload.npz"
if not os.path.exists(cache):
    model, trajs, seeds, _ = bench.build_workload("cfg3_funnel1024_p64_k1000_j6", 0, 1)
    offsets, X, G = pf.Engine.pack(trajs)
    np.savez(cache, offsets=offsets, X=X, G=G, seeds=np.concatenate(seeds))
d = np.load(cache)
model = pf.Funnel(1024)
eng = pf.Engine(1024, model.family, model.blob, 6, 1000, 0, two_pass=(mode != "lean"), materialize_all=(mode == "m"))
eng.upload(d["offsets"], d["X"], d["G"], d["seeds"])
for _ in range(5):
    eng.run()
    eng.sync()
print(eng.timings())
