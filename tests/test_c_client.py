"""The drop-in boundary is a C ABI: a plain C99 program (tests/c_client/abi_client.c — no Python, no torch,
only include/pfb200.h) drives libpfb200.so the way the Julia shim does through ccall.

CPU: the header compiles as strict C99, the client links against the library, and without a CUDA
device the engine refuses to come up (exit code 3: the product path has no CPU fallback).
GPU: the client's dump (its inputs and every output) is replayed through the ctypes binding and the
two callers must agree bit for bit — and the ELBOs must match the oracle."""
import os
import pathlib
import struct
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "c_client" / "abi_client.c"
LIBDIR = ROOT / "pathfinder_b200"


def _build(tmp_path):
    exe = tmp_path / "abi_client"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(SRC), "-o",
           str(exe), f"-L{LIBDIR}", "-lpfb200", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def _have_lib():
    return (LIBDIR / "libpfb200.so").exists()


@pytest.mark.skipif(not _have_lib(), reason="libpfb200.so has not been built (python -c 'import __graft_entry__ as g; g.build()')")
def test_plain_c_client_builds_and_fails_loudly_without_a_device(tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the GPU test below runs the client")
    r = subprocess.run([str(exe), str(tmp_path / "dump.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "pfb_create failed" in r.stderr and "no CPU fallback" in r.stderr
    assert not (tmp_path / "dump.bin").exists()


def _read_dump(path):
    b = path.read_bytes()
    pos = 0

    def take(dtype, count):
        nonlocal pos
        a = np.frombuffer(b, dtype=dtype, count=count, offset=pos).copy()
        pos += a.nbytes
        return a

    n, P, K, J, ndraws, T, U, inputs_only = struct.unpack_from("8i", b, 0)
    pos = 32
    d = {"n": n, "P": P, "K": K, "J": J, "ndraws": ndraws, "T": T, "U": U}
    d["offsets"] = take(np.int64, P + 1)
    d["X"] = take(np.float64, n * T).reshape(n, T, order="F")
    d["G"] = take(np.float64, n * T).reshape(n, T, order="F")
    d["seeds"] = take(np.uint64, U)
    if inputs_only:
        assert pos == len(b)
        return d
    d["elbo"] = take(np.float64, U)
    d["se"] = take(np.float64, U)
    d["best_iter"] = take(np.int64, P)
    d["success"] = take(np.int32, P)
    d["weights"] = take(np.float64, K * P)
    d["pareto_k"] = float(take(np.float64, 1)[0])
    d["tail_len"] = int(take(np.int64, 1)[0])
    d["inds"] = take(np.int64, ndraws)
    d["ids"] = take(np.int64, ndraws)
    d["draws"] = take(np.float64, n * ndraws).reshape(n, ndraws, order="F")
    assert pos == len(b)
    return d


def _oracle_elbos(d):
    from oracle import pf_oracle as O

    off, out, u = d["offsets"], [], 0
    for p in range(d["P"]):
        Xp, Gp = d["X"][:, off[p]:off[p + 1]], d["G"][:, off[p]:off[p + 1]]
        mus, Hs, _ = O.fit_mvnormals(Xp, Gp, history_length=d["J"])
        for l in range(1, Xp.shape[1]):
            un = np.asarray(O.contract_normals(int(d["seeds"][u]), d["n"], d["K"]))
            x, lq = O.rand_and_logpdf(un, mus[:, l], Hs[l])
            out.append(np.mean(-0.5 * np.sum(x * x, axis=0) - lq))
            u += 1
    assert u == d["U"]
    return np.array(out)


@pytest.mark.skipif(not _have_lib(), reason="libpfb200.so has not been built")
def test_c_client_inputs_give_finite_oracle_elbos(tmp_path):
    # the client's synthetic traces (--inputs-only needs no device) are a sane ELBO problem for the oracle:
    # every iteration has a finite ELBO, so the GPU leg's success / best_iter assertions are meaningful
    exe = _build(tmp_path)
    r = subprocess.run([str(exe), "--inputs-only", str(tmp_path / "in.bin")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    d = _read_dump(tmp_path / "in.bin")
    assert (d["n"], d["P"], d["K"], d["J"]) == (24, 3, 64, 6) and list(np.diff(d["offsets"])) == [6, 4, 8]
    assert np.array_equal(d["G"][:, 0] != 0, np.ones(24, bool))
    e = _oracle_elbos(d)
    assert e.shape == (15,) and np.isfinite(e).all()


@pytest.mark.gpu
def test_plain_c_client_agrees_with_the_ctypes_binding_and_the_oracle(tmp_path):
    import pathfinder_b200 as pf

    exe = _build(tmp_path)
    dump = tmp_path / "dump.bin"
    r = subprocess.run([str(exe), str(dump)], capture_output=True, text=True, timeout=300,
                       env={**os.environ, "CUDA_VISIBLE_DEVICES": os.environ.get("CUDA_VISIBLE_DEVICES", "0")})
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    d = _read_dump(dump)
    assert d["success"].all() and (d["best_iter"] >= 1).all()

    # the same inputs through the Python binding: identical outputs, bit for bit
    eng = pf.Engine(d["n"], 0, None, d["J"], d["K"], 0)
    res = eng.elbo_batch(d["offsets"], d["X"], d["G"], d["seeds"], draws=False)
    assert np.array_equal(res.elbo, d["elbo"]) and np.array_equal(res.elbo_se, d["se"])
    assert np.array_equal(res.best_iter, d["best_iter"]) and np.array_equal(res.success, d["success"])
    rr = eng.psis_resample(2024, d["ndraws"], True)
    assert np.array_equal(rr["weights"], d["weights"]) and rr["pareto_k"] == d["pareto_k"]
    assert rr["tail_len"] == d["tail_len"]
    assert np.array_equal(rr["inds"], d["inds"]) and np.array_equal(rr["ids"], d["ids"])
    assert np.array_equal(rr["draws"], d["draws"])
    assert np.array_equal(d["ids"], (d["inds"] - 1) // d["K"] + 1)  # cld(ind, K), src/resample.jl:70
    eng.close()

    # and the C caller's ELBOs against the oracle (engine RNG contract, iso-normal target)
    np.testing.assert_allclose(d["elbo"], _oracle_elbos(d), rtol=1e-6, atol=1e-6)
