#!/usr/bin/env python3
"""bench.py — ELBO Monte-Carlo samples/s of the hot path (BASELINE.json metric).

One "step" = one pass of the whole hot path over one batch: K1 history scan -> K2 Woodbury
build -> K3 fused sampling/logq/logp -> K4 ELBO reduce + argmax -> K5 best-iteration draws ->
(NCCL all-gather of the per-path log ratios when N > 1) -> K6 PSIS -> K7 resample + gather.

Workload (config.workload): BASELINE config 3 — multipathfinder on a 1024-dim Neal's funnel,
64 paths per GPU, K = 1000 draws per iteration, history 6; trajectories come from the host
L-BFGS (master seed 20261017, per-path seed = master + global path index, init U[-10, 10]).
Paths shard across ranks with no data-path collective until the PSIS pool (weak scaling).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement on all host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MASTER_SEED = 20261017
METRIC = "elbo_mc_samples_per_sec"
UNIT = "samples/s"

CONFIGS = {
    # name: (model kind, n, paths per GPU, K, J, init_scale, ndraws)
    "cfg3_funnel1024_p64_k1000_j6": ("funnel", 1024, 64, 1000, 6, 10.0, 1000),
    "cfg2_funnel100_p8_k1000_j6": ("funnel", 100, 8, 1000, 6, 10.0, 1000),
    # hierarchical logistic regression, 254 features + (log tau, b0), 2048 rows
    "cfg4_hlogistic256_p32_k2000_j6": ("hlogistic", 256, 32, 2000, 6, 2.0, 1000),
    # 4096-dim correlated Gaussian, history 10 (16 paths per GPU = 128 on 8 GPUs)
    "cfg5_dense4096_p16_k500_j10": ("dense", 4096, 16, 500, 10, 2.0, 1000),
}


def make_model(kind, n):
    import pathfinder_b200 as pf

    if kind == "funnel":
        return pf.Funnel(n), 3.0 * n
    if kind == "hlogistic":
        nobs, p = 2048, n - 2
        Xm = np.random.default_rng(4).normal(size=(nobs, p))
        beta = np.random.default_rng(5).normal(size=p) * 0.5
        y = (np.random.default_rng(55).random(nobs) < 1.0 / (1.0 + np.exp(-(Xm @ beta)))).astype(np.float64)
        return pf.HierLogistic(Xm, y), 2.0 * nobs * p + 30.0 * nobs
    if kind == "dense":
        rng = np.random.default_rng(6)
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        lam = rng.random(n) * 0.95 + 0.05
        prec = (Q / lam) @ Q.T  # Sigma = Q diag(lam) Q' (test/test_utils.jl:7-10)
        return pf.DenseNormal(np.random.default_rng(7).normal(size=n), 0.5 * (prec + prec.T)), 2.0 * n * n
    raise ValueError(kind)


def build_workload(name, rank, world, strong=False, assign="lpt"):
    """Trajectories of this rank's paths.  N > 1: every rank optimises all paths (seeded, identical
    everywhere, untimed set-up) and takes its share of an LPT assignment on the iteration counts
    (work is proportional to L_p).  Weak scaling: P paths per rank (world x P in total);
    strong scaling (BASELINE config 3 as stated: 64 paths over 1 -> 8 GPUs): P paths in total."""
    import pathfinder_b200 as pf

    kind, n, P, K, J, scale, ndraws = CONFIGS[name]
    if strong:
        if P % world:
            raise SystemExit(f"--scaling strong needs the path count {P} to be a multiple of --gpus")
        P = P // world
    model, model_flops = make_model(kind, n)

    def one(gp):
        rng = np.random.default_rng(MASTER_SEED + gp)
        x0 = (rng.random(n) * 2.0 - 1.0) * scale
        tr = pf.optimize_with_trace(model, x0, J, 1000)
        return (tr.points, tr.gradients), rng.integers(0, 2**64, size=len(tr) - 1, dtype=np.uint64)

    if world == 1 or assign == "static":
        # static: rank r owns the contiguous block of paths [r P, (r + 1) P) and optimises only those (targets
        # whose paths all take about the same number of iterations; the host L-BFGS of config 5 costs ~2 s a path)
        mine = list(range(rank * P, (rank + 1) * P))
        paths = {gp: one(gp) for gp in mine}
    else:
        paths = {gp: one(gp) for gp in range(world * P)}
        L = np.array([paths[gp][0][0].shape[1] - 1 for gp in range(world * P)])
        loads, owned = np.zeros(world), [[] for _ in range(world)]
        for gp in np.argsort(-L, kind="stable"):
            r = min((q for q in range(world) if len(owned[q]) < P), key=lambda q: (loads[q], q))
            loads[r] += L[gp]
            owned[r].append(int(gp))
        mine = sorted(owned[rank])
    trajs = [paths[gp][0] for gp in mine]
    seeds = [paths[gp][1] for gp in mine]
    return model, trajs, seeds, (n, P, K, J, ndraws, model_flops)


def bench_config(name, n, P, K, J, ndraws):
    """The `config` object, identical for both arms (the driver compares them key by key)."""
    return {"workload": name, "n": n, "paths_per_gpu": P, "K": K, "history": J, "ndraws": ndraws}


def algorithmic_flops(trajs, n, K, J, model_flops):
    """SURVEY §8(d) mode-F count per sample: 2n (|u|^2) + k^2 (Vc') + 4nk (Vh'u, Vh w) +
    2k^2 (T) + 2n (scale, shift) + F_model (3n funnel, 2n^2 dense normal, ...), k = 2 min(l, J)."""
    total = 0.0
    for X, _ in trajs:
        for l in range(1, X.shape[1]):
            k = 2 * min(l, J)
            total += K * (2 * n + k * k + 4 * n * k + 2 * k * k + 2 * n + model_flops)
    return total


def algorithmic_bytes_mode_m(trajs, n, K, J):
    """SURVEY §8(d) mode-M bytes: 8n + 24 per sample + [16n + 16n(2J+2)] per unit."""
    U = sum(X.shape[1] - 1 for X, _ in trajs)
    return U * (K * (8 * n + 24) + 16 * n + 16 * n * (2 * J + 2))


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_logp(kind, model):
    from oracle import pf_oracle as O

    if kind == "funnel":
        return O.logp_funnel
    if kind == "dense":
        return O.make_logp_dense_gaussian(model.mean, model.prec)
    return O.make_logp_hier_logistic(model.X, model.y)


def oracle_elbo_stage(model_n, trajs, seeds, K, J, budget_s, max_paths=None, logp_fn=None):
    """The CPU restatement (oracle) of the ELBO stage on a bounded sample; returns
    (samples done, seconds)."""
    from oracle import pf_oracle as O

    logp_fn = logp_fn or O.logp_funnel
    done, t0 = 0, time.perf_counter()
    for p, (X, G) in enumerate(trajs):
        if max_paths is not None and p >= max_paths:
            break
        mus, Hs, _ = O.fit_mvnormals(X, G, history_length=J)
        L = X.shape[1] - 1
        for l in range(1, L + 1):
            u = O.contract_normals(int(seeds[p][l - 1]), model_n, K)
            O.elbo_and_samples(u, logp_fn, mus[:, l], Hs[l])
            done += K
            if time.perf_counter() - t0 > budget_s:
                return done, time.perf_counter() - t0
    return done, time.perf_counter() - t0


def _ref_worker(args):
    from threadpoolctl import threadpool_limits

    n, X, G, sd, K, J, budget, kind = args
    model, _ = make_model(kind, n)
    with threadpool_limits(limits=1):  # one BLAS thread per worker process, one process per core
        return oracle_elbo_stage(n, [(X, G)], [sd], K, J, budget, logp_fn=oracle_logp(kind, model))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia cannot run here) on
    all host cores, paths across processes like ntasks = nthreads (src/multipath.jl:190).  A step is the
    ELBO stage of the workload's own P paths — all of them when that fits the per-step budget, else the
    largest LPT-ordered prefix that does (calibrated in the first warm-up step) — handed to the workers
    one path at a time, longest first, so that no core idles behind a straggler."""
    if rank != 0:
        return
    import multiprocessing as mp

    name = args.config
    model, trajs, seeds, (n, P, K, J, ndraws, model_flops) = build_workload(name, 0, 1)
    P = len(trajs)
    cores = os.cpu_count() or 1
    budget_s = 4.0
    L = np.array([X.shape[1] - 1 for X, _ in trajs])
    order = [int(p) for p in np.argsort(-L, kind="stable")]  # longest paths first
    ctx = mp.get_context("fork")
    vals, times = [], []
    sample = list(order)
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            jobs = [(n, trajs[p][0], trajs[p][1], seeds[p], K, J, 1e9, CONFIGS[name][0]) for p in sample]
            t0 = time.perf_counter()
            out = list(pool.imap_unordered(_ref_worker, jobs, chunksize=1))
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                vals.append(sum(o[0] for o in out) / dt)
                times.append(dt)
            if step == 0 and dt > budget_s and len(sample) > cores:
                # bounded sample: every other path of the LPT order keeps the length mix of the workload
                keep = max(cores, int(len(sample) * budget_s / dt))
                sample = [sample[int(round(i * (len(sample) - 1) / max(1, keep - 1)))] for i in range(keep)]
                sample = sorted(set(sample), key=lambda p: -L[p])
    v = float(np.mean(vals))
    units = int(sum(L[p] for p in sample))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(name, n, P, K, J, ndraws),
        "details": {"mode": "reference CPU algorithm (every iteration's draws materialised, as src/elbo.jl:19)",
                    "parallelism": f"paths over {cores} host processes, longest first (imap_unordered)",
                    "paths_per_step": len(sample), "units_per_step": units},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"ELBO stage (fit_mvnormals + maximize_elbo) of the oracle port on {len(sample)} of "
                                   f"the workload's {P} paths ({units} units x {K} draws) per step; Julia is not "
                                   f"installed so the reference itself cannot run"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3_funnel1024_p64_k1000_j6", choices=list(CONFIGS))
    ap.add_argument("--cpu-budget", type=float, default=0.0,
                    help="seconds of CPU work for cpu_baseline (0: the whole ELBO stage, ~1 min, so that the "
                         "multipathfinder wall-clock ratio divides two measurements)")
    ap.add_argument("--no-mode-m", action="store_true", help="skip the secondary mode-M (materialise) timing")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle timing (profiling runs)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's paths PER GPU; strong: the config's paths in total, split over the GPUs")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-wall", action="store_true", help="skip the multipathfinder() wall-clock legs")
    ap.add_argument("--assign", default=None, choices=["lpt", "static"],
                    help="N > 1: LPT-balanced on iteration counts (default for the funnels; every rank optimises every "
                         "path during set-up) or contiguous blocks (default for the GEMM-shaped targets)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import pathfinder_b200 as pf
    from pathfinder_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists for the product path)")
    torch.cuda.set_device(local_rank)
    # several ranks on one host: keep each rank (and the page-locked staging it first-touches) on its GPU's NUMA node
    numa = None
    if world > 1 and not args.no_numa:
        from pathfinder_b200._numa import bind_to_device_numa
        numa = bind_to_device_numa(local_rank)
    # NCCL prints its version banner on stdout (NCCL_DEBUG=VERSION): keep stdout = the one JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    name = args.config
    assign = args.assign or ("lpt" if CONFIGS[name][0] == "funnel" else "static")
    model, trajs, seeds, (n, P, K, J, ndraws, model_flops) = build_workload(name, rank, world,
                                                                            strong=args.scaling == "strong", assign=assign)
    P = len(trajs)
    U = sum(X.shape[1] - 1 for X, _ in trajs)
    offsets, X, G = pf.Engine.pack(trajs)
    seeds_cat = np.concatenate(seeds)

    lib = _lib.load()
    import ctypes as C
    peak = C.c_double(0.0)
    lib.pfb_measure_fp64_dmma_tflops(local_rank, 5, C.byref(peak))  # FP64 tensor-core (DMMA) peak
    fp64_peak = peak.value
    peak2 = C.c_double(0.0)
    lib.pfb_measure_fp64_fma_tflops(local_rank, 5, C.byref(peak2))  # DFMA peak, reported alongside

    eng = pf.Engine(n, model.family, model.blob, J, K, local_rank)
    eng.upload(offsets, X, G, seeds_cat)  # inputs resident in HBM before the timed region
    view = eng.device_view()
    ext = torch.cuda.ExternalStream(view.stream, device=local_rank)

    def wrap(ptr, numel):
        class _A:
            __cuda_array_interface__ = {"shape": (numel,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_A(), device=f"cuda:{local_rank}")

    counts = [P] * world  # every rank owns P paths (LPT-balanced); rank order = run order of the pool
    if world > 1:
        eng.comm_init(None)  # the library's own NCCL communicator (pfb_comm_init), id broadcast by the group

    resample_seed = MASTER_SEED

    step_bufs = {"r": None}  # caller-owned, page-locked result buffers reused across steps

    def resample(want_weights, into):
        # the product call: pfb_psis_resample on one GPU, pfb_pool_exchange_resample on several (all-gather
        # of the per-draw log densities, PSIS + index draw replicated, owned columns regenerated, sum-reduce
        # — all on the engine stream behind the C ABI, no host synchronisation in between)
        if world == 1:
            return eng.psis_resample(resample_seed, ndraws, True, into=into)
        return eng.pool_exchange_resample(counts, resample_seed, ndraws, True, want_weights=want_weights, into=into)

    def step():
        eng.run()
        # device-resident timing: at N > 1 the N-sized weight vectors stay on the device (the e2e leg
        # below brings them to the host)
        r = resample(world == 1, step_bufs["r"])
        if step_bufs["r"] is None:
            eng.pin(r.get("log_weights"), r.get("weights"), r["draws"], r["inds"], r["ids"])
            step_bufs["r"] = r
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k3_ms, stage_ms = [], []
    with torch.cuda.stream(ext):
        ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        tm = eng.timings()
        k3_ms.append(tm["k3"]); stage_ms.append(tm)
    with torch.cuda.stream(ext):
        ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    tt = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    uu = torch.tensor([float(U)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(uu, op=dist.ReduceOp.SUM)
    max_ms = float(tt.item())
    total_units = float(uu.item())
    value = total_units * K * args.steps / (max_ms * 1e-3)

    # ---- e2e: the public call with HOST buffers (pinned), copies inside the timed region ---------
    _keep = []

    def pinned(a):
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory()
        _keep.append(t)
        v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        return v
    XT, GT = pinned(np.ascontiguousarray(X.T)), pinned(np.ascontiguousarray(G.T))  # F-order == C-order of .T
    Xp, Gp = XT.T, GT.T
    sp, op_ = pinned(seeds_cat), pinned(offsets)

    # caller-owned output buffers, reused across steps and page-locked once (what a Julia caller does
    # with preallocated arrays through the C ABI)
    e2e_bufs = {"res": None, "r": None}

    def e2e_step():
        res = eng.elbo_batch(op_, Xp, Gp, sp, draws=False, fit=True, into=e2e_bufs["res"])
        r = resample(True, e2e_bufs["r"])
        if e2e_bufs["res"] is None:
            eng.pin(*pf.Engine.result_arrays(res))
            e2e_bufs["res"] = res
            eng.pin(r["log_weights"], r["weights"], r["draws"], r["inds"], r["ids"])
            e2e_bufs["r"] = r
        return res, r

    e2e_step()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res, r = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_units * K * args.steps / float(te.item())
    KP = eng.KP
    h2d = 2 * X.nbytes + seeds_cat.nbytes + offsets.nbytes
    Npool = world * K * P
    d2h = 16 * U + 20 * P + (n * P * 2 + n * KP * P + 2 * KP * KP * P + P) * 8 + 4 * P \
        + 16 * Npool + 16 * ndraws + 8 * n * ndraws

    # ---- roofline of the dominant kernel (K3; its Q-apply runs on the FP64 tensor cores) -----------
    k3_avg_ms = float(np.mean(k3_ms))
    flops = algorithmic_flops(trajs, n, K, J, model_flops)
    ach_tf = flops / (k3_avg_ms * 1e-3) / 1e12
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "k3_traffic.json")))
        if tj.get("workload") == name:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach_tf / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
                "kernel": "pfb_k3_elbo_sample (lean: FP64 DMMA Q-apply, Philox/ziggurat normals; single pass for the "
                          "diagonal-quadratic families, K3 + the fused FP64 tensor-core GEMM K8g for the GEMM-shaped ones)",
                "kernel_ms": k3_avg_ms, "algorithmic_flops_per_launch": flops,
                "peak_source": "measured live: pfb_measure_fp64_dmma_tflops (mma.sync m8n8k4 f64 chains); "
                               "MEASURED_PEAKS.json has bf16 and HBM figures only",
                "dfma_peak_tflops": peak2.value,
                "note": "FP64-pipe bound: per 8-row x 16-draw block 12 DMMA (192 clk of pipe) + 20 scalar FP64 (~66 clk) + one "
                        "Philox4x32-7 call (12 IMAD.WIDE) for the lane's four 32-bit ziggurat variates; the pipe is 86 % busy "
                        "inside the hot loop, 26 % of the instructions are outside it (profiles/r2_k3_queue_ncu.md, "
                        "profiles/r2_k3_experiments.md); HBM is idle: see DESIGN.md section 4",
                "k3_share_of_step": k3_avg_ms * args.steps / max_ms}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(name, n, P, K, J, ndraws),
        "details": {"units_per_gpu_rank0": U, "mode": "F (lean: per-draw logp/logq only; best-iteration draws "
                    "regenerated on demand)", "l2": "per-step working set (factor records %.0f MB) exceeds the "
                    "126 MB L2" % (U * n * 16 * 8 / 1e6),
                    "parallelism": f"paths sharded over {world} GPU(s)" + (
                        ", LPT-balanced on iteration counts; pool exchange = pfb_pool_exchange_resample (NCCL behind "
                        "the C ABI)" if world > 1 else "")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        # K1..K5 as counted by the engine + the PSIS stage's own kernels (K6a..K6g, K7; CUB's sort
        # launches inside K6 are library kernels and not counted)
        "gpu_launches": int((stage_ms[-1]["launches"] + 10) * args.steps),
        "clocks": clocks,
        "roofline": roofline,
        "stage_ms": {k: float(np.mean([s[k] for s in stage_ms])) for k in ("k1", "k2", "k3", "k4", "k5", "total")},
        "wall_s_timed_region": wall,
    }
    if world > 1:
        line["numa_bind_rank0"] = numa  # {"node", "cpus"} when the rank was bound to its GPU's NUMA node, else null

    # ---- whole multipathfinder() call, trajectories included (north_star's wall-clock target) ----------
    # The DEFAULT call (maxiters = 1000, ntries = 1) with the host optimiser (the reference's split:
    # sequential L-BFGS on the CPU, everything after it on the GPU) and with the device optimiser K0,
    # on a caller-owned (warm) engine; best of 3 after one warm-up call.  The CPU side is MEASURED
    # below (cpu_baseline leg): the single-thread port of the same call.
    mpf = None
    if rank == 0 and world == 1 and not args.no_wall and model.family in (0, 1, 2, 3, 4):
        mpf = {"paths": P, "K": K, "ndraws": ndraws, "maxiters": 1000, "ntries": 1}
        for opt in ("host", "device"):
            engw = pf.Engine(n, model.family, model.blob, J, K, local_rank)
            walls, units_w = [], 0
            for rep_ in range(4):
                t0 = time.perf_counter()
                rw = pf.multipathfinder(model, ndraws, nruns=P, ndraws_elbo=K, rng=np.random.default_rng(MASTER_SEED),
                                        init_scale=CONFIGS[name][5], maxiters=1000, optimizer=opt, engine=engw,
                                        ntries=1, history_length=J)
                walls.append(time.perf_counter() - t0)
                units_w = sum(len(pr.elbo_estimates) for pr in rw.pathfinder_results)
                rw = None  # drop the result: its (never accessed) per-path draws need not leave the device
            mpf[opt] = {"ours_s": float(min(walls[1:])), "ours_first_call_s": float(walls[0]), "units": int(units_w)}
            if opt == "device":
                mpf[opt]["k0_ms"] = float(engw.lbfgs_ms())
            engw.close()
        line["multipathfinder_wall"] = mpf

    if rank == 0 and world == 1 and not args.no_mode_m and float(U) * n * K * 8 < 40e9:
        # secondary accounting: reference-faithful mode M (every iteration's draws written to HBM)
        engm = pf.Engine(n, model.family, model.blob, J, K, local_rank, materialize_all=True)
        engm.upload(offsets, X, G, seeds_cat)
        for _ in range(2):
            engm.run(); engm.sync()
        ms = []
        for _ in range(max(3, args.steps // 2)):
            engm.run(); engm.sync()
            ms.append(engm.timings()["k3"])
        by = algorithmic_bytes_mode_m(trajs, n, K, J)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        gbs = by / (float(np.mean(ms)) * 1e-3) / 1e9
        line["roofline_mode_m"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": gbs / hbm_peak, "traffic": None, "kernel_ms": float(np.mean(ms)),
                                   "value_samples_per_s": U * K / (float(np.mean(ms)) * 1e-3),
                                   "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}
        engm.close()

    if rank == 0:
        # ---- cpu_baseline: the oracle port on one host core ----------------------------------------
        # With the wall-clock leg: the WHOLE single-thread port of the call is measured (host L-BFGS of
        # every path + ELBO stage of every unit + PSIS / resample), so the speed-up below divides two
        # measurements; --cpu-budget bounds the ELBO-stage sample otherwise.
        from threadpoolctl import threadpool_limits

        full_cpu = mpf is not None and not args.no_cpu_baseline and args.cpu_budget <= 0
        if args.no_cpu_baseline or world > 1:  # (the CPU baseline is timed at N = 1 only)
            done, secs = 0, 1.0
        else:
            with threadpool_limits(limits=1):
                done, secs = oracle_elbo_stage(n, trajs, seeds, K, J, 1e9 if full_cpu else max(args.cpu_budget, 15.0),
                                               logp_fn=oracle_logp(CONFIGS[name][0], model))
        line["cpu_baseline"] = {"value": (done / secs) if done else None, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": (f"oracle ELBO stage on {'all' if done == U * K else 'the first'} "
                                           f"{done // K} (path, iteration) units of this workload, {secs:.1f} s, "
                                           f"single thread") if done else "not timed in this run (N > 1 or --no-cpu-baseline)"}
        if mpf is not None and done > 0:
            from oracle import psis as OP

            with threadpool_limits(limits=1):
                t0 = time.perf_counter()
                for p in range(P):  # the same host L-BFGS the GPU arm's host-optimiser call runs
                    pf.optimize_with_trace(model, (np.random.default_rng(MASTER_SEED + p).random(n) * 2 - 1)
                                           * CONFIGS[name][5], J, 1000)
                cpu_lbfgs_s = time.perf_counter() - t0
                t0 = time.perf_counter()
                lr = np.random.default_rng(1).normal(size=P * K)
                pr_ = OP.psis(lr)
                OP.resample_indices(7, pr_["weights"], lr.size, ndraws)
                cpu_psis_s = time.perf_counter() - t0
            cpu_elbo_s = secs if done == U * K else U * K / (done / secs)
            cpu_s = cpu_lbfgs_s + cpu_elbo_s + cpu_psis_s
            mpf["cpu_port_single_thread_s"] = float(cpu_s)
            mpf["cpu_port_parts_s"] = {"lbfgs": float(cpu_lbfgs_s), "elbo_stage": float(cpu_elbo_s),
                                       "psis_resample": float(cpu_psis_s),
                                       "elbo_stage_measured": bool(done == U * K)}
            for opt in ("host", "device"):
                mpf[opt]["speedup_vs_cpu_port"] = float(cpu_s / mpf[opt]["ours_s"])
                # (the device optimiser does not stop where SciPy does: it may run many more iterations, i.e.
                # units; the per-sample rates make the two comparable)
                mpf[opt]["elbo_samples_per_s_wall"] = float(mpf[opt]["units"] * K / mpf[opt]["ours_s"])
            mpf["cpu_port_elbo_samples_per_s_wall"] = float(U * K / cpu_s)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
