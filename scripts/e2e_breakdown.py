#!/usr/bin/env python3
"""Where the end-to-end step's time goes (config 3): upload / run / download / PSIS, and the fused call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, pathfinder_b200 as pf

name = "cfg3_funnel1024_p64_k1000_j6"
model, trajs, seeds, (n, P, K, J, ndraws, _) = bench.build_workload(name, 0, 1)
offsets, X, G = pf.Engine.pack(trajs)
sd = np.concatenate(seeds)
keep = []
def pinned(a):
    t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory(); keep.append(t)
    v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape); v[...] = a; return v
XT, GT = pinned(np.ascontiguousarray(X.T)), pinned(np.ascontiguousarray(G.T))
Xp, Gp, sp, op_ = XT.T, GT.T, pinned(sd), pinned(offsets)
eng = pf.Engine.for_model(model, J, K, 0)
def T(f, reps=10):
    f(); eng.sync()
    t = time.perf_counter()
    for _ in range(reps): f()
    eng.sync()
    return (time.perf_counter() - t) / reps * 1e3
res = eng.elbo_batch(op_, Xp, Gp, sp, draws=False, fit=True)
r = eng.psis_resample(7, ndraws, True)
eng.pin(*pf.Engine.result_arrays(res)); eng.pin(r["log_weights"], r["weights"], r["draws"], r["inds"], r["ids"])
print("upload            %.3f ms" % T(lambda: eng.upload(op_, Xp, Gp, sp)))
print("run (K1..K5)      %.3f ms" % T(lambda: (eng.run(), eng.sync())))
print("download fit      %.3f ms" % T(lambda: eng.download(draws=False, fit=True, into=res)))
print("psis+resample     %.3f ms" % T(lambda: eng.psis_resample(7, ndraws, True, into=r)))
print("separate sequence %.3f ms" % T(lambda: (eng.upload(op_, Xp, Gp, sp), eng.run(), eng.download(draws=False, fit=True, into=res), eng.psis_resample(7, ndraws, True, into=r))))
print("fused elbo_batch + psis %.3f ms" % T(lambda: (eng.elbo_batch(op_, Xp, Gp, sp, draws=False, fit=True, into=res), eng.psis_resample(7, ndraws, True, into=r))))
