#!/usr/bin/env python3
"""Single-process multi-GPU check of the C ABI (what a Julia caller without MPI does): one engine per GPU in
ONE process, pfb_comm_init_all, every engine runs its shard of the runs, ONE pfb_pool_exchange_resample_all
call — compared with the same pool resampled on a single GPU.

    python scripts/multi_inproc_check.py [ndev]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathfinder_b200 as pf  # noqa: E402
from pathfinder_b200 import _lib  # noqa: E402
from pathfinder_b200._lib import pfb_resample_out  # noqa: E402
from tests.helpers import synthetic_trajectory  # noqa: E402

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n, K, J, ndraws = 48, 64, 6, 100
model = pf.Funnel(n)
shards = [[synthetic_trajectory(n, 3 + (7 * r + j) % 5, 100 * r + j, scale=0.4) for j in range(2 + r)] for r in range(ndev)]
lib = _lib.load()
engs = [pf.Engine.for_model(model, J, K, r) for r in range(ndev)]
hs = (C.c_void_p * ndev)(*[e.h for e in engs])
rc = lib.pfb_comm_init_all(hs, ndev)
assert rc == 0, (rc, lib.pfb_last_error(engs[0].h))
seeds = []
for r, (e, trajs) in enumerate(zip(engs, shards)):
    offsets, X, G = pf.Engine.pack(trajs)
    sd = np.random.default_rng(r).integers(0, 2**64, size=int(offsets[-1]) - len(trajs), dtype=np.uint64)
    seeds.append(sd)
    e.upload(offsets, X, G, sd)
    e.run()
ppr = np.array([len(s) for s in shards], dtype=np.int32)
N = int(ppr.sum()) * K
outs = (pfb_resample_out * ndev)()
keep = []
for r in range(ndev):
    o, res = engs[r]._resample_out(N, ndraws, True, True)
    outs[r] = o
    keep.append(res)
rc = lib.pfb_pool_exchange_resample_all(hs, ndev, ppr.ctypes.data_as(C.c_void_p), C.c_uint64(11), ndraws, 1, 1, outs)
assert rc == 0, (rc, lib.pfb_last_error(engs[0].h))
for r in range(1, ndev):
    for k in ("inds", "ids", "weights", "draws"):
        assert np.array_equal(keep[0][k], keep[r][k]), (r, k)
# the same pool on one GPU
one = pf.Engine.for_model(model, J, K, 0)
trajs = [t for s in shards for t in s]
offsets, X, G = pf.Engine.pack(trajs)
one.upload(offsets, X, G, np.concatenate(seeds))
one.run()
ref = one.psis_resample(11, ndraws, True)
for k in ("inds", "ids", "weights", "draws"):
    assert np.array_equal(ref[k], keep[0][k]), k
print(f"single-process exchange over {ndev} GPUs == single-GPU result (inds, ids, weights, draws); "
      f"k-hat {float(keep[0]['pareto_k'][0]):.3f}")
for e in engs + [one]:
    e.close()
