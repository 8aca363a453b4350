"""Multi-GPU plumbing: one process per GPU over ``torch.distributed`` (NCCL on the GPUs, gloo in
the CPU tests).

On the GPUs the exchange itself lives behind the C ABI (``pfb_comm_init`` +
``pfb_pool_exchange_resample``, NCCL inside libpfb200.so; ``Engine.comm_init`` /
``Engine.pool_exchange_resample``) — this module keeps the sharding rules both sides use and
``pooled_resample``, the same exchange written over ``torch.distributed`` tensors, which the CPU
tests run under gloo (world size 2) to pin the host-side logic: ragged shards, ranks without runs,
pool order.

The reference is single-process (OhMyThreads tasks over runs, src/multipath.jl:190-208).  Here the
paths shard across ranks; K1..K5 run with no communication, and the only exchange is the PSIS pool
(lean variant): all-gather the per-draw log densities (16 B per pool draw), run PSIS
and the index draw replicated on every rank (deterministic kernels + counter-based RNG => the
same indices everywhere), then sum-reduce the ``ndraws`` selected columns, each contributed by
the rank that owns it.  The pool keeps the reference's order: draw-fastest, component-slowest
with components in the original run order (src/multipath.jl:217, test/resample.jl:81-88), because
ranks own contiguous blocks of runs.
"""
from __future__ import annotations

import numpy as np


def shard_range(nruns: int, rank: int, world: int):
    """Contiguous, balanced block of runs owned by `rank` (keeps the global component order)."""
    return nruns * rank // world, nruns * (rank + 1) // world


def shard_counts(nruns: int, world: int):
    return [shard_range(nruns, r, world)[1] - shard_range(nruns, r, world)[0] for r in range(world)]


def is_distributed(group=None) -> bool:
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def allgather_ragged(x, counts, group=None):
    """All-gather 1-D tensors whose lengths differ per rank (`counts[r]` elements from rank r):
    pad to the longest, one all_gather_into_tensor, strip the padding."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    m = max(counts) if counts else 0
    send = x.new_zeros(m)
    send[: x.numel()] = x
    recv = x.new_empty(world * m)
    dist.all_gather_into_tensor(recv, send, group=group)
    return torch.cat([recv[r * m: r * m + counts[r]] for r in range(world)])


def pooled_resample(local_logp, local_logq, local_pool, K_run, nruns, psis_resample_fn, seed, ndraws,
                    importance=True, group=None, device="cpu"):
    """The cross-rank part of `_compute_psis_result` + `_resample` (src/multipath.jl:220-225).

    local_logp / local_logq: [K_run * P_local] float64 (draw-fastest) of this rank's runs;
    local_pool: [n, K_run * P_local] the matching draws (NumPy, F-order) or None;
    psis_resample_fn(log_ratios or None, N) -> dict with 1-based "inds" (+ weights, pareto_k, ...):
        the single-rank PSIS + index draw (the engine's kernel K6/K7; the oracle in the CPU tests).
    Returns that dict plus "ids" and, if a pool was given, "draws" [n, ndraws] on every rank."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = [c * K_run for c in shard_counts(nruns, world)]
    lo = sum(counts[:rank])
    N = sum(counts)
    logr = None
    if importance:
        lp = torch.as_tensor(np.ascontiguousarray(local_logp), dtype=torch.float64, device=device)
        lq = torch.as_tensor(np.ascontiguousarray(local_logq), dtype=torch.float64, device=device)
        g = allgather_ragged(torch.cat([lp, lq]), [2 * c for c in counts], group)
        # per rank the payload is [logp | logq]; rebuild the two global vectors
        off, gp, gq = 0, [], []
        for c in counts:
            gp.append(g[off: off + c]); gq.append(g[off + c: off + 2 * c]); off += 2 * c
        logr = (torch.cat(gp) - torch.cat(gq)).cpu().numpy()
    r = dict(psis_resample_fn(logr, N))
    inds = np.asarray(r["inds"], dtype=np.int64)
    r["ids"] = -(-inds // K_run)  # cld(ind, K_run), src/resample.jl:70
    if local_pool is not None:
        n = local_pool.shape[0]
        mine = (inds > lo) & (inds <= lo + counts[rank])
        out = np.zeros((n, ndraws), order="F")
        out[:, mine] = local_pool[:, inds[mine] - 1 - lo]
        t = torch.as_tensor(out.T.copy(), device=device)  # [ndraws, n] contiguous
        dist.all_reduce(t, group=group)
        r["draws"] = np.asfortranarray(t.cpu().numpy().T)
    return r
