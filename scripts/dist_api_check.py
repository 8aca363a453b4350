#!/usr/bin/env python3
"""Multi-GPU check of the public API under NCCL: torchrun --nproc-per-node N scripts/dist_api_check.py
Every rank calls pf.multipathfinder with the same arguments and seed.  Checked: (1) draws, ids, PSIS
weights, k-hat and sample indices are identical on all ranks; (2) they are IDENTICAL to the same call
run by one process on one GPU (runs keep their order across ranks, kernels are deterministic);
(3) every resampled column is a pool column of the run its component id names.  Cases: both
optimisers, ragged shards, fewer runs than ranks (a rank without runs), top-up draws beyond
ndraws_elbo (pool assembled on the host, pfb_pool_set), uniform resampling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pathfinder_b200 as pf

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, K, ndraws = 64, 200, 300
model = pf.Funnel(n)
eng = pf.Engine.for_model(model, 6, K, local)  # caller-owned: per-path draws stay on the device until looked at
cases = [dict(nruns=10, optimizer="device"), dict(nruns=10, optimizer="host"), dict(nruns=world + 1, optimizer="device"),
         dict(nruns=max(1, world - 1), optimizer="device"), dict(nruns=6, optimizer="device", ndraws_per_run=K + 50),
         dict(nruns=7, optimizer="device", importance=False)]
for case in cases:
    kw = dict(ndraws_elbo=K, init_scale=4.0, maxiters=25, device=local)
    kw.update(case)
    nruns = kw["nruns"]
    r = pf.multipathfinder(model, ndraws, rng=np.random.default_rng(5), engine=eng, **kw)
    assert r.draws.shape == (n, ndraws) and r.draw_component_ids.min() >= 1 and r.draw_component_ids.max() <= nruns
    w = r.psis_result.weights if r.psis_result is not None else np.zeros(1)
    kh = r.psis_result.pareto_shape if r.psis_result is not None else 0.0
    h = torch.tensor([float(np.nansum(r.draws)), float(r.draw_component_ids.sum()), float(np.nansum(w)), float(kh),
                      float(r.sample_inds.sum())], dtype=torch.float64, device=f"cuda:{local}")
    g = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(g, h)
    assert all(torch.equal(g[0], x) for x in g), ("ranks disagree", case, [x.tolist() for x in g])
    # the rank's own runs: columns attributed to them exist in their pools
    lo, hi = nruns * rank // world, nruns * (rank + 1) // world
    assert len(r.pathfinder_results) == hi - lo
    for j, pr in enumerate(r.pathfinder_results):
        cid = lo + j + 1
        cols = r.draws[:, r.draw_component_ids == cid]
        pool = pr.draws
        for c in cols.T:
            assert np.any(np.all(pool == c[:, None], axis=0)), "resampled column not in the owner's pool"
    if rank == 0:
        one = pf.multipathfinder(model, ndraws, rng=np.random.default_rng(5), group=False, **kw)
        assert np.array_equal(one.sample_inds, r.sample_inds), ("indices differ from the single-GPU run", case)
        assert np.array_equal(one.draws, r.draws) and np.array_equal(one.draw_component_ids, r.draw_component_ids)
        if r.psis_result is not None:
            assert np.array_equal(one.psis_result.weights, r.psis_result.weights)
        print(f"{case}: {world} ranks agree with each other and with the single-GPU run; sum(weights)="
              f"{h[2].item():.12f} k-hat={h[3].item():.3f}", flush=True)
    dist.barrier()
eng.close()
dist.destroy_process_group()
