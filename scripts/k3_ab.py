#!/usr/bin/env python3
"""A/B timing of K3 builds on the bench workload (config 3): for every library given, K1..K4 stage
times of pfb_batch_run with inputs resident (median of the timed steps).

    python scripts/k3_ab.py pathfinder_b200/libpfb200.so pathfinder_b200/libpfb200_X.so ...
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import sys, json, numpy as np
sys.path.insert(0, %r)
import bench, pathfinder_b200 as pf
d = np.load(%r, allow_pickle=True)
offsets, X, G, seeds = d["offsets"], d["X"], d["G"], d["seeds"]
n, K, J = 1024, 1000, 6
model = pf.Funnel(n)
eng = pf.Engine(n, model.family, model.blob, J, K, 0, two_pass=%s, materialize_all=%s)
eng.upload(offsets, X, G, seeds)
ms = []
for i in range(9):
    eng.run(); eng.sync()
    if i >= 3: ms.append(eng.timings())
print(json.dumps({k: float(np.median([m[k] for m in ms])) for k in ("k1", "k2", "k3", "k4", "total")}))
"""


def main():
    import numpy as np

    import bench

    cache = "/tmp/k3_ab_workload.npz"
    if not os.path.exists(cache):
        import pathfinder_b200 as pf

        model, trajs, seeds, _ = bench.build_workload("cfg3_funnel1024_p64_k1000_j6", 0, 1)
        offsets, X, G = pf.Engine.pack(trajs)
        np.savez(cache, offsets=offsets, X=X, G=G, seeds=np.concatenate(seeds))
    mode = os.environ.get("K3_AB_MODE", "lean")
    for lib in sys.argv[1:]:
        env = dict(os.environ, PFB200_LIB=os.path.abspath(lib))
        code = CHILD % (ROOT, cache, "True" if mode != "lean" else "False", "True" if mode == "m" else "False")
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(os.path.basename(lib), mode, out.stdout.strip() or out.stderr[-400:], flush=True)


if __name__ == "__main__":
    main()
