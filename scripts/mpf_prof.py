import sys, os, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np
import pathfinder_b200 as pf
n, P, K = 1024, 64, 1000
model = pf.Funnel(n)
eng = pf.Engine.for_model(model, 6, K, 0)
def call():
    return pf.multipathfinder(model, 1000, nruns=P, ndraws_elbo=K, rng=np.random.default_rng(20261017), init_scale=10.0,
                              maxiters=64, optimizer="device", engine=eng, ntries=1)
call(); call()
t=time.perf_counter(); call(); print("wall", time.perf_counter()-t)
pr = cProfile.Profile(); pr.enable(); call(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
