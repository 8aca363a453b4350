import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pathfinder_b200 as pf
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def log(*a):
    print(f"[r{rank} {time.time()%1000:.2f}]", *a, flush=True)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg up")
n, K = 32, 64
model = pf.Funnel(n)
eng = pf.Engine.for_model(model, 6, K, local)
log("engine up")
eng.comm_init(None)
log("comm up")
from tests.helpers import synthetic_trajectory
P = 2 + rank
trajs = [synthetic_trajectory(n, 4, 10 * rank + j) for j in range(P)]
offsets, X, G = pf.Engine.pack(trajs)
U = int(offsets[-1]) - P
eng.upload(offsets, X, G, np.arange(U, dtype=np.uint64))
eng.run(); eng.sync()
log("ran")
counts = [2 + r for r in range(world)]
r = eng.pool_exchange_resample(counts, 7, 20, True)
log("exchange ok", r["inds"][:5], float(r["draws"].sum()))
counts = [3] * world
if rank == 0:
    trajs = trajs + [synthetic_trajectory(n, 4, 99)]
    offsets, X, G = pf.Engine.pack(trajs)
    U = int(offsets[-1]) - 3
    eng.upload(offsets, X, G, np.arange(U, dtype=np.uint64)); eng.run(); eng.sync()
r = eng.pool_exchange_resample(counts, 7, 20, True)
log("equal exchange ok", r["inds"][:5], float(r["draws"].sum()))
eng.close()
dist.destroy_process_group()
log("done")
