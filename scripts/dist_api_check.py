#!/usr/bin/env python3
"""Multi-GPU check of the public API under NCCL: torchrun --nproc-per-node N scripts/dist_api_check.py
Every rank calls pf.multipathfinder with the same arguments and seed; the result (draws, ids, PSIS
weights) must be identical on all ranks and every resampled column must be a pool column of the
run its component id names."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pathfinder_b200 as pf

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, nruns, K, ndraws = 64, 10, 200, 300
model = pf.Funnel(n)
for optimizer in ("device", "host"):
    r = pf.multipathfinder(model, ndraws, nruns=nruns, ndraws_elbo=K, rng=np.random.default_rng(5), init_scale=4.0,
                           maxiters=25, optimizer=optimizer, device=local)
    assert r.draws.shape == (n, ndraws) and r.draw_component_ids.min() >= 1 and r.draw_component_ids.max() <= nruns
    h = torch.tensor([float(np.nansum(r.draws)), float(r.draw_component_ids.sum()), float(np.nansum(r.psis_result.weights)),
                      float(r.psis_result.pareto_shape)], dtype=torch.float64, device=f"cuda:{local}")
    g = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(g, h)
    assert all(torch.equal(g[0], x) for x in g), "ranks disagree"
    # the rank's own runs: columns attributed to them exist in their pools
    lo, hi = nruns * rank // world, nruns * (rank + 1) // world
    for j, pr in enumerate(r.pathfinder_results):
        cid = lo + j + 1
        cols = r.draws[:, r.draw_component_ids == cid]
        pool = pr.draws
        for c in cols.T:
            assert np.any(np.all(pool == c[:, None], axis=0)), "resampled column not in the owner's pool"
    if rank == 0:
        print(f"optimizer={optimizer}: {world} ranks agree; sum(weights)={h[2].item():.12f} k-hat={h[3].item():.3f}", flush=True)
dist.destroy_process_group()
