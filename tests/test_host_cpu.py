"""CPU tests (no GPU) of the host side: the C ABI surface, the loud failure without a device, the
trajectory producer, batch packing, and the multi-rank pool exchange over gloo (world size 2)."""
import os
import re
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


# ---- C ABI ------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    """include/pfb200.h is the contract: every function it declares is exported by libpfb200.so
    and bound by the ctypes layer (no compute calls here)."""
    from pathfinder_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "pfb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pfb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layouts_match_header():
    import ctypes as C

    from pathfinder_b200 import _lib

    assert C.sizeof(_lib.pfb_config) == 32
    assert C.sizeof(_lib.pfb_elbo_out) == 18 * C.sizeof(C.c_void_p)
    assert C.sizeof(_lib.pfb_resample_out) == 7 * C.sizeof(C.c_void_p)
    assert C.sizeof(_lib.pfb_device_view) == 5 * C.sizeof(C.c_void_p) + 4 * 8
    assert C.sizeof(_lib.pfb_lbfgs_opts) == 24


@pytest.mark.skipif(_have_gpu(), reason="checks the no-device behaviour")
def test_no_device_fails_loudly():
    """There is no CPU fallback: without a CUDA device the engine raises, with the CUDA error text."""
    import pathfinder_b200 as pf

    with pytest.raises(pf.PfbError) as ei:
        pf.Engine(4, pf.PFB_MODEL_ISONORMAL, None, 6, 5, 0)
    assert ei.value.code != 0
    with pytest.raises(pf.PfbError):
        pf.pathfinder(pf.IsoNormal(3), rng=np.random.default_rng(0))


def test_argument_errors_do_not_need_a_device():
    import ctypes as C

    from pathfinder_b200 import _lib

    lib = _lib.load()
    h = C.c_void_p()
    bad = _lib.pfb_config(0, 0, 5, 0, 0, 0, 1e-12)  # history_length = 0
    assert lib.pfb_create(C.byref(h), C.byref(bad)) == -1
    bad = _lib.pfb_config(0, 65, 5, 0, 0, 0, 1e-12)  # history_length > 64 unsupported (13..64: generic kernels)
    assert lib.pfb_create(C.byref(h), C.byref(bad)) == -3
    assert lib.pfb_create(None, None) == -1
    assert lib.pfb_batch_run(None) == -1 and lib.pfb_destroy(None) == 0


# ---- host logic ---------------------------------------------------------------------------------
def test_optimize_with_trace_records_points_and_gradients():
    """src/optimize.jl:86-108: every iterate with its log density and gradient of the LOG density;
    the run ends at a non-finite value."""
    import pathfinder_b200 as pf

    model = pf.Funnel(6)
    x0 = np.array([1.0, 0.5, -0.5, 0.2, 0.1, -0.3])
    tr = pf.optimize_with_trace(model, x0, 6, 50)
    assert tr.points.shape == tr.gradients.shape and tr.points.shape[0] == 6
    assert np.array_equal(tr.points[:, 0], x0)
    for l in range(len(tr)):
        assert np.isclose(tr.log_densities[l], model.logp(tr.points[:, l]))
        np.testing.assert_allclose(tr.gradients[:, l], model.grad(tr.points[:, l]))
    assert tr.log_densities[-1] >= tr.log_densities[0]

    class Bad(pf.IsoNormal):
        def logp(self, x):
            return float("nan")

    # non-finite at the initial point: the callback pushes the point, then stops (src/optimize.jl:94-105)
    bad = pf.optimize_with_trace(Bad(3), np.ones(3))
    assert len(bad) == 1 and np.isnan(bad.log_densities[0]) and np.array_equal(bad.points[:, 0], np.ones(3))


def test_model_gradients_match_finite_differences():
    import pathfinder_b200 as pf

    rng = np.random.default_rng(0)
    for model in (pf.IsoNormal(5), pf.Funnel(5), pf.DiagNormal(rng.normal(size=5), rng.random(5) + 0.5)):
        x = rng.normal(size=5)
        g = model.grad(x)
        for i in range(5):
            e = np.zeros(5); e[i] = 1e-6
            assert abs((model.logp(x + e) - model.logp(x - e)) / 2e-6 - g[i]) < 1e-5 * max(1, abs(g[i]))


def test_pack_layout():
    import pathfinder_b200 as pf

    a = (np.arange(6.0).reshape(2, 3), -np.arange(6.0).reshape(2, 3))
    b = (np.ones((2, 1)), np.zeros((2, 1)))
    off, X, G = pf.Engine.pack([a, b])
    assert off.tolist() == [0, 3, 4] and X.flags.f_contiguous and X.shape == (2, 4)
    assert np.array_equal(X[:, :3], a[0]) and np.array_equal(G[:, 3:], b[1])


def test_shard_ranges_cover_all_runs():
    from pathfinder_b200 import distributed as D

    for nruns in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = D.shard_range(nruns, r, world)
                got += list(range(lo, hi))
            assert got == list(range(nruns))
            c = D.shard_counts(nruns, world)
            assert sum(c) == nruns and max(c) - min(c) <= 1


# ---- multi-rank pool exchange over gloo ----------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nruns, K_run, n, ndraws, importance, q):
    import torch.distributed as dist

    from oracle import psis as OP
    from pathfinder_b200 import distributed as D

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(123)  # every rank builds the same global pool, keeps its block
    logp = rng.standard_t(4, size=nruns * K_run)
    logq = rng.normal(size=nruns * K_run)
    pool = np.asfortranarray(rng.normal(size=(n, nruns * K_run)))
    lo, hi = D.shard_range(nruns, rank, world)
    sl = slice(lo * K_run, hi * K_run)

    def psis_fn(logr, N):  # stand-in for the engine's K6/K7 (the oracle computes the same contract)
        if logr is None:
            return dict(inds=OP.resample_indices(5, None, N, ndraws))
        res = OP.psis(logr)
        return dict(inds=OP.resample_indices(5, res["weights"], N, ndraws), weights=res["weights"],
                    log_weights=res["log_weights"], pareto_k=res["pareto_k"], tail_len=res["tail_length"])

    r = D.pooled_resample(logp[sl], logq[sl], pool[:, sl], K_run, nruns, psis_fn, 5, ndraws, importance)
    # single-process answer on the whole pool
    ref = psis_fn(logp - logq if importance else None, nruns * K_run)
    ok = (np.array_equal(r["inds"], ref["inds"]) and np.array_equal(r["draws"], pool[:, ref["inds"] - 1])
          and np.array_equal(r["ids"], -(-ref["inds"] // K_run)))
    if importance:
        ok = ok and np.array_equal(r["weights"], ref["weights"])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nruns,importance", [(4, True), (5, True), (3, False), (1, True)])
def test_pool_exchange_world_size_2_gloo(nruns, importance):
    """Two ranks, ragged shards (5 runs -> 2 + 3): the gathered log ratios, the replicated index
    draw and the owner-contributed columns equal the single-process result bit for bit."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nruns, 10, 3, 25, importance, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_bench_lpt_assignment_is_a_balanced_partition():
    """bench.py (N > 1): every rank computes the same LPT assignment of the world x P paths on their
    iteration counts; the shares are disjoint, equally sized and balanced."""
    import bench

    name, world = "cfg2_funnel100_p8_k1000_j6", 4
    shares, units = [], []
    for rank in range(world):
        _, trajs, seeds, (n, P, K, J, nd, _) = bench.build_workload(name, rank, world)
        assert len(trajs) == P and all(s.size == X.shape[1] - 1 for (X, _), s in zip(trajs, seeds))
        shares.append({X[:, 0].tobytes() for X, _ in trajs})        # a path is identified by its initial point
        units.append(sum(X.shape[1] - 1 for X, _ in trajs))
    assert len(set().union(*shares)) == world * P                    # disjoint, complete
    assert max(units) <= 1.1 * (sum(units) / world)
    _, t1, _, _ = bench.build_workload(name, 0, 1)
    assert len(t1) == 8


def test_product_code_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under pathfinder_b200/ (Python or CUDA) may import,
    link or include it, and libpfb200.so must not depend on the oracle's shared objects."""
    import subprocess

    pkg = os.path.join(ROOT, "pathfinder_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), (dirpath, f)
                for line in txt.splitlines():
                    code = line.split("//")[0].split("#", 1)[0] if not line.lstrip().startswith("#include") else line
                    assert not (("#include" in code or f == "Makefile") and "oracle" in code), (dirpath, f, line)
    so = os.path.join(pkg, "libpfb200.so")
    if os.path.exists(so):
        deps = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
        assert "pforacle" not in deps


def test_numa_cpulist_parse_and_bind_is_harmless():
    # pathfinder_b200/_numa.py: host plumbing of the multi-rank bench (bind a rank to its GPU's NUMA node);
    # without a GPU nothing may change and nothing may raise
    import os

    from pathfinder_b200 import _numa

    assert _numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _numa._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert _numa.bind_to_device_numa(0) is None
    assert os.sched_getaffinity(0) == before


def _run_bench(args, env=None, timeout=300):
    import json
    import os
    import pathlib
    import subprocess
    import sys

    root = pathlib.Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout, env={**os.environ, **(env or {})}, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    return [json.loads(ln) for ln in lines]


def test_reference_arm_prints_the_contract_line_and_only_rank_0_works():
    """bench.py --impl reference: the CPU algorithm (oracle port) on the host cores with the SAME metric,
    unit and config keys as the product arm, a measured ms_per_step, e2e = value with zero copies; under
    torchrun only rank 0 runs it, the other ranks exit 0 silently."""
    cfg = "cfg2_funnel100_p8_k1000_j6"
    (line,) = _run_bench(["--impl", "reference", "--config", cfg, "--steps", "2", "--warmup", "1"])
    assert line["impl"] == "reference" and line["metric"] == "elbo_mc_samples_per_sec" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"] == {"workload": cfg, "n": 100, "paths_per_gpu": 8, "K": 1000, "history": 6, "ndraws": 1000}
    assert line["value"] > 0 and line["ms_per_step"] > 0
    # value = samples of a step / its measured duration
    units = line["details"]["units_per_step"]
    assert abs(line["value"] - units * 1000 / (line["ms_per_step"] / 1e3)) / line["value"] < 0.2
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "paths" in cb["sample"]
    # a non-zero rank of a multi-rank launch does nothing and prints nothing
    assert _run_bench(["--impl", "reference", "--config", cfg, "--gpus", "2", "--steps", "1", "--warmup", "0"],
                      env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_julia_shim_ccalls_match_the_header():
    """julia/PathfinderB200.jl cannot be executed here (no Julia): at least every ccall in it must name a
    function that include/pfb200.h declares, with the same number of arguments and a pointer / scalar
    pattern that agrees with the C prototype."""
    import pathlib
    import re

    root = pathlib.Path(__file__).resolve().parents[1]
    hdr = re.sub(r"/\*.*?\*/", " ", (root / "include" / "pfb200.h").read_text(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\s*\*|size_t)\s+(pfb_\w+)\s*\(([^;{}]*?)\)\s*;", hdr, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",")] if m.group(2).strip() not in ("", "void") else []
        protos[m.group(1)] = args
    assert len(protos) >= 35

    jl = (root / "julia" / "PathfinderB200.jl").read_text()
    calls = re.findall(r"ccall\(\(:(pfb_\w+), LIB\),\s*(\w+),\s*\((.*?)\)\s*,", jl, flags=re.S)
    assert len(calls) >= 15 and "LIB[]" not in jl
    # the binding shown in INTEGRATION.md is held to the same standard
    doc_calls = re.findall(r"ccall\(\(:(pfb_\w+), LIB\),\s*(\w+),\s*\((.*?)\)\s*,", (root / "INTEGRATION.md").read_text(),
                           flags=re.S)
    assert len(doc_calls) >= 4
    calls = calls + doc_calls
    seen = set()
    for name, ret, types in calls:
        assert name in protos, f"{name} is not declared in pfb200.h"
        seen.add(name)
        jt = [t.strip() for t in types.replace("\n", " ").split(",") if t.strip()]
        ct = protos[name]
        assert len(jt) == len(ct), (name, jt, ct)
        assert ret in ("Cint", "Cstring"), (name, ret)
        for j, c in zip(jt, ct):
            c_is_ptr = "*" in c or "pfb_handle" in c or "pfb_logp_callback" in c
            j_is_ptr = j.startswith(("Ptr{", "Ref{")) or j == "Cstring"
            assert c_is_ptr == j_is_ptr, (name, j, c)
            if not c_is_ptr:
                width = {"Cint": "int", "Int32": "int32_t", "Int64": "int64_t", "UInt64": "uint64_t", "Csize_t": "size_t",
                         "Cdouble": "double", "Float64": "double"}[j]
                assert re.search(rf"\b{width}\b", c), (name, j, c)
    # the calls a drop-in needs are all there
    assert {"pfb_create", "pfb_destroy", "pfb_register_model", "pfb_elbo_batch", "pfb_psis_resample",
            "pfb_set_fallback_seeds", "pfb_unit_fits", "pfb_comm_init", "pfb_pool_exchange_resample",
            "pfb_pool_exchange_resample_all", "pfb_register_host_model", "pfb_lbfgs_batch"} <= seen


def test_julia_shim_structs_mirror_the_header_field_for_field():
    import pathlib
    import re

    root = pathlib.Path(__file__).resolve().parents[1]
    hdr = re.sub(r"/\*.*?\*/", " ", (root / "include" / "pfb200.h").read_text(), flags=re.S)
    jl = (root / "julia" / "PathfinderB200.jl").read_text()
    ctype = {"Int32": "int32_t", "Int64": "int64_t", "Float64": "double", "Ptr{Float64}": "double*",
             "Ptr{Int64}": "int64_t*", "Ptr{Int32}": "int32_t*"}
    pairs = {"PfbConfig": "pfb_config", "PfbElboOut": "pfb_elbo_out", "PfbResampleOut": "pfb_resample_out",
             "PfbResampleOutC": "pfb_resample_out", "PfbLbfgsOpts": "pfb_lbfgs_opts"}
    for jname, cname in pairs.items():
        body = re.search(rf"typedef struct\s*\{{([^}}]*)\}}\s*{cname}\s*;", hdr).group(1)
        cfields = [(re.sub(r"\s+", "", t), n) for t, n in re.findall(r"([\w ]+?\*?)\s*(\w+)\s*;", body)]
        jbody = re.search(rf"struct {jname}\n(.*?)\nend", jl, flags=re.S).group(1)
        jfields = [(n, t) for n, t in re.findall(r"^\s*(\w+)::([\w{{}}]+)\s*$", jbody, flags=re.M)]
        assert len(cfields) == len(jfields) > 0, (jname, cfields, jfields)
        for (ct, cn), (jn, jt) in zip(cfields, jfields):
            assert cn == jn and ctype[jt] == ct, (jname, (ct, cn), (jn, jt))


def test_reference_citations_point_at_existing_lines():
    """Every `src/...jl:LINE` style citation in the header, the docs and the oracle names a file of the
    reference that exists and is at least that long (checked where /root/reference is present)."""
    import pathlib
    import re

    ref = pathlib.Path("/root/reference")
    if not ref.is_dir():
        pytest.skip("/root/reference is not present on this machine")
    root = pathlib.Path(__file__).resolve().parents[1]
    files = [root / "include" / "pfb200.h", root / "DESIGN.md", root / "INTEGRATION.md", root / "julia" / "PathfinderB200.jl"]
    files += sorted((root / "oracle").glob("*.py")) + sorted((root / "pathfinder_b200").glob("*.py"))
    files += sorted((root / "oracle").glob("*.c*")) + [root / "bench.py", root / "README.md"]
    files += [f for ext in ("*.cu", "*.cuh", "*.h") for f in sorted((root / "pathfinder_b200" / "csrc").glob(ext))
              if f.name != "pf_zig_tables.h"]
    pat = re.compile(r"\b((?:src|test|ext|docs/src)/[\w/.\-]+\.(?:jl|md)):(\d+)(?:-(\d+))?")
    nlines, checked = {}, 0
    for f in files:
        for m in pat.finditer(f.read_text()):
            path = ref / m.group(1)
            assert path.is_file(), f"{f.name}: cites {m.group(1)}, which the reference does not have"
            if path not in nlines:
                nlines[path] = len(path.read_text().splitlines())
            last = int(m.group(3) or m.group(2))
            assert 1 <= int(m.group(2)) <= last <= nlines[path], f"{f.name}: {m.group(0)} (file has {nlines[path]} lines)"
            checked += 1
    assert checked > 100
