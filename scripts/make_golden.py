#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference (Julia) cannot run in the build container and ships no golden vectors for this path
(SURVEY §8c), so these files pin the ENGINE'S OWN CONTRACT (Philox/ziggurat normals, PSIS
arithmetic, resample index stream) and the oracle's restatement of the reference algorithm on
fixed inputs: any later change of either shows up as a diff against the committed numbers.
`tests/test_golden.py` checks the oracle (CPU) and the CUDA path (GPU) against them.

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pf_oracle as O  # noqa: E402
from oracle import psis as OP  # noqa: E402
from tests.helpers import synthetic_trajectory  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. the normal-variate contract
    seeds = np.array([0, 1, 0xDEADBEEFCAFEF00D, 2**64 - 1], dtype=np.uint64)
    normals = np.stack([np.asarray(O.contract_normals(int(s), 9, 6)) for s in seeds])
    np.savez(os.path.join(OUT, "normals_contract.npz"), seeds=seeds, normals=normals)

    # 2. one ELBO batch: two paths of an isotropic-normal target, n = 16, history 6, K = 48
    n, K, J = 16, 48, 6
    trajs = [synthetic_trajectory(n, L, 900 + L) for L in (3, 7)]
    rng = np.random.default_rng(17)
    sd = [rng.integers(0, 2**64, size=X.shape[1] - 1, dtype=np.uint64) for X, _ in trajs]
    out = dict(n=n, K=K, J=J)
    for p, (X, G) in enumerate(trajs):
        mus, Hs, rej = O.fit_mvnormals(X, G, history_length=J)
        lopt, ests = O.maximize_elbo(sd[p], O.logp_isonormal, mus, Hs, K)
        out.update({f"X{p}": X, f"G{p}": G, f"seeds{p}": sd[p], f"rejected{p}": rej, f"lopt{p}": lopt,
                    f"elbo{p}": np.array([e["value"] for e in ests]),
                    f"se{p}": np.array([e["std_err"] for e in ests]),
                    f"logp{p}": np.stack([e["logp"] for e in ests], 1),
                    f"logq{p}": np.stack([e["logq"] for e in ests], 1),
                    f"draws{p}": ests[lopt - 1]["draws"], f"mu{p}": mus[:, lopt],
                    f"logdet{p}": Hs[lopt].logdet()})
    np.savez(os.path.join(OUT, "elbo_batch.npz"), **out)

    # 3. PSIS + resample on fixed log ratios
    rng = np.random.default_rng(99)
    lr = rng.standard_t(3, size=3000) * 1.5
    res = OP.psis(lr)
    inds = OP.resample_indices(4242, res["weights"], lr.size, 64)
    uinds = OP.resample_indices(4242, None, lr.size, 64)
    np.savez(os.path.join(OUT, "psis_resample.npz"), log_ratios=lr, log_weights=res["log_weights"],
             weights=res["weights"], pareto_k=res["pareto_k"], tail_length=res["tail_length"], seed=4242,
             inds=inds, uniform_inds=uinds,
             norep_inds=OP.resample_indices_norep(4242, res["log_weights"], lr.size, 64),
             norep_uniform_inds=OP.resample_indices_norep(4242, None, lr.size, 64))

    # 4. the L-BFGS trajectory contract (pf_lbfgs.h): funnel and independent normals, small n
    from oracle import lbfgs as OL

    rng = np.random.default_rng(5)
    x0f = rng.uniform(-4, 4, size=12)
    Xf, FXf, Gf, stf, nevf = OL.lbfgs_path(OL.FAMILY_FUNNEL, x0f, 6, 25)
    mean, sd = rng.normal(size=9), rng.uniform(0.2, 5.0, size=9)
    x0d = rng.uniform(-2, 2, size=9)
    Xd, FXd, Gd, std, nevd = OL.lbfgs_path(OL.FAMILY_DIAGNORMAL, x0d, 6, 1000, mean=mean, sd=sd)
    np.savez(os.path.join(OUT, "lbfgs_traces.npz"), x0f=x0f, Xf=Xf, FXf=FXf, Gf=Gf, stf=stf, nevf=nevf,
             mean=mean, sd=sd, x0d=x0d, Xd=Xd, FXd=FXd, Gd=Gd, std=std, nevd=nevd)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
