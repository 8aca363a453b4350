"""Engine: NumPy-facing wrapper of one libpfb200 handle (one GPU).

Host buffers in, host buffers out — exactly what a Julia shim does through ``ccall`` — plus
the split upload / run / download calls used by ``bench.py`` to time the kernels with inputs
resident in HBM.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import pfb_config, pfb_device_view, pfb_elbo_out, pfb_lbfgs_opts, pfb_resample_out


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class ElboBatchResult:
    """Per-batch outputs (unit order: path-major, iteration-minor)."""

    offsets: np.ndarray          # [P+1] point offsets
    elbo: np.ndarray             # [U]
    elbo_se: np.ndarray          # [U]
    best_iter: np.ndarray        # [P] 1-based fit_iteration, 0 = none
    success: np.ndarray          # [P] bool
    n_rejected: np.ndarray       # [P]
    draws: np.ndarray | None = None        # [n, K, P] (F-order) best-iteration draws
    draws_logp: np.ndarray | None = None   # [K, P]
    draws_logq: np.ndarray | None = None   # [K, P]
    logp: np.ndarray | None = None         # [K, U]
    logq: np.ndarray | None = None         # [K, U]
    fit: dict | None = None                # best-iteration factors (reference WoodburyPDMat form)
    all_draws: np.ndarray | None = None    # [n, K, U] when materialize_all

    def unit_slice(self, p):
        o = self.offsets
        return slice(int(o[p] - p), int(o[p + 1] - p - 1))

    # lazy pool (download(draws="lazy")): the best-iteration draws stay on the device until looked at
    _lazy_engine = None

    def fetch_draws(self):
        """Bring the (device-resident) best-iteration draws of this batch to the host, once."""
        if self.draws is None and self._lazy_engine is not None:
            eng = self._lazy_engine()
            self._lazy_engine = None
            if eng is None or eng.h is None:
                raise RuntimeError("the engine holding these draws is gone")
            self.draws, self.draws_logp, self.draws_logq = eng._pool_download()
        return self.draws


class Engine:
    def __init__(self, n, model_family, model_blob=None, history_length=6, ndraws_elbo=5,
                 device=0, materialize_all=False, eps=1e-12, two_pass=False, host_logp=None):
        self.lib = _lib.load()
        self.n = int(n)
        self.K = int(ndraws_elbo)
        self.J = int(history_length)
        cfg = pfb_config(int(device), self.J, self.K, int(bool(materialize_all)), int(bool(two_pass)), 0, float(eps))
        h = C.c_void_p()
        rc = self.lib.pfb_create(C.byref(h), C.byref(cfg))
        if rc != 0:
            raise _lib.PfbError(rc, (self.lib.pfb_last_error(None) or b"").decode())
        self.h = h
        self.KP = self.lib.pfb_kp(self.h)
        self.materialize_all = bool(materialize_all)
        self._cb = None
        self._cb_error = None
        if int(model_family) == _lib.PFB_MODEL_HOSTCALLBACK:
            if host_logp is None:
                raise ValueError("a host-callback model needs host_logp(X[n, m]) -> logp[m]")
            self._register_host(host_logp)
        else:
            blob = None if model_blob is None else np.ascontiguousarray(model_blob, dtype=np.float64)
            _lib.check(self.h, self.lib.pfb_register_model(
                self.h, int(model_family), self.n, _ptr(blob), 0 if blob is None else blob.size))
        self._P = 0
        self._U = 0
        self._offsets = None
        self._poolK = self.K
        self._poolP = None  # paths in the device pool when it differs from the batch's (pool_set)
        self._comm = None
        self._pending = []  # weakrefs of results whose draws are still on the device

    @classmethod
    def for_model(cls, model, history_length=6, ndraws_elbo=5, device=0, **kw):
        """Engine for a model object of models.py (registered family, or HostModel)."""
        return cls(model.n, model.family, getattr(model, "blob", None), history_length, ndraws_elbo, device,
                   host_logp=getattr(model, "logp_batch", None), **kw)

    def _register_host(self, host_logp):
        """Row f2: the target density is a host function (the reference's `logp` closure,
        src/elbo.jl:15).  The C callback wraps the pinned staging buffer as an n x m F-order array
        without copying; an exception inside the callback is re-raised by run()."""
        n = self.n

        def cb(_user, xptr, n_, m, out):
            try:
                X = np.ctypeslib.as_array(xptr, shape=(int(m), int(n_))).T
                res = np.ctypeslib.as_array(out, shape=(int(m),))
                res[:] = np.asarray(host_logp(X), dtype=np.float64).reshape(int(m))
            except BaseException as e:  # noqa: BLE001 — must not unwind through C
                self._cb_error = e
                np.ctypeslib.as_array(out, shape=(int(m),))[:] = np.nan

        self._cb = _lib.pfb_logp_callback(cb)  # keep the trampoline alive as long as the engine
        _lib.check(self.h, self.lib.pfb_register_host_model(self.h, n, self._cb, None))

    def _raise_cb_error(self):
        if self._cb_error is not None:
            e, self._cb_error = self._cb_error, None
            raise e

    def _pool_download(self):
        P, K = self._P, self._poolK
        draws = np.empty((self.n, K, P), order="F")
        lp = np.empty((K, P), order="F")
        lq = np.empty((K, P), order="F")
        _lib.check(self.h, self.lib.pfb_pool_download(self.h, 0, P, _ptr(draws), _ptr(lp), _ptr(lq)))
        return draws, lp, lq

    def _flush_pending(self):
        """The device pool is about to be overwritten / freed: results that still point at it get
        their draws now."""
        pend, self._pending = self._pending, []
        for ref in pend:
            res = ref()
            if res is not None and res.draws is None and res._lazy_engine is not None:
                res.fetch_draws()

    def close(self):
        if getattr(self, "h", None) and getattr(self, "_pending", None):
            try:
                self._flush_pending()
            except Exception:
                pass
        if getattr(self, "h", None):
            self.lib.pfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ELBO stage ------------------------------------------------------------------------
    @staticmethod
    def pack(trajectories, n=None):
        """trajectories: list of (points [n, L+1], gradients [n, L+1]) -> offsets, X, G (F-order).
        n: the dimension, needed only for an empty list (a rank that owns no run)."""
        P = len(trajectories)
        offsets = np.zeros(P + 1, dtype=np.int64)
        for p, (x, _) in enumerate(trajectories):
            offsets[p + 1] = offsets[p] + x.shape[1]
        n = trajectories[0][0].shape[0] if P else int(n or 0)
        X = np.empty((n, int(offsets[-1])), dtype=np.float64, order="F")
        G = np.empty_like(X)
        for p, (x, g) in enumerate(trajectories):
            X[:, offsets[p]:offsets[p + 1]] = x
            G[:, offsets[p]:offsets[p + 1]] = g
        return offsets, X, G

    def set_fallback_seeds(self, seeds):
        """One UInt64 seed per path of the NEXT batch: a failed path's draws are fresh draws from the fit
        of its best iteration with this seed (src/singlepath.jl:224-228; pfb_set_fallback_seeds)."""
        sd = np.ascontiguousarray(seeds, dtype=np.uint64)
        _lib.check(self.h, self.lib.pfb_set_fallback_seeds(self.h, int(sd.size), _ptr(sd)))

    def upload(self, offsets, X, G, seeds, normals=None):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        P = offsets.size - 1
        U = int(offsets[-1]) - P
        X = np.asfortranarray(X, dtype=np.float64)
        G = np.asfortranarray(G, dtype=np.float64)
        if X.shape != (self.n, int(offsets[-1])) or G.shape != X.shape:
            raise ValueError("positions / gradients must be n x offsets[-1]")
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        if seeds.size != U:
            raise ValueError("need one seed per (path, iteration)")
        if normals is not None:
            normals = np.asfortranarray(normals, dtype=np.float64)
            if normals.shape != (self.n, self.K, U):
                raise ValueError("normals must be n x K x U")
        self._flush_pending()
        _lib.check(self.h, self.lib.pfb_batch_upload(self.h, self.n, P, _ptr(offsets), _ptr(X), _ptr(G),
                                                     _ptr(seeds), _ptr(normals)))
        self._P, self._U, self._offsets = P, U, offsets.copy()

    # ---- device L-BFGS (K0) ----------------------------------------------
    def lbfgs_batch(self, x0, maxiters=1000, max_points=None, gtol=1e-8, ftol=1e-14):
        """Run the L-BFGS trajectories of P paths on the device (one CTA per path; replaces
        optimize_with_trace, src/optimize.jl:35-59).  x0: [n, P].  Returns (npoints [P], status [P],
        nevals [P]); the traces stay resident (batch_from_lbfgs / lbfgs_download)."""
        x0 = np.asfortranarray(x0, dtype=np.float64)
        if x0.ndim != 2 or x0.shape[0] != self.n:
            raise ValueError("x0 must be n x P")
        P = x0.shape[1]
        o = pfb_lbfgs_opts(int(maxiters), int(max_points or maxiters + 1), float(gtol), float(ftol))
        npts = np.zeros(P, dtype=np.int64)
        st = np.zeros(P, dtype=np.int32)
        nev = np.zeros(P, dtype=np.int32)
        _lib.check(self.h, self.lib.pfb_lbfgs_batch(self.h, self.n, P, _ptr(x0), C.byref(o), _ptr(npts), _ptr(st),
                                                    _ptr(nev)))
        self._lb_npts = npts
        return npts, st, nev

    def batch_from_lbfgs(self, seeds):
        """Make the device-resident traces of the last lbfgs_batch the current batch (what upload()
        does from host buffers); seeds: one per (path, iteration)."""
        npts = self._lb_npts
        P = npts.size
        offsets = np.zeros(P + 1, dtype=np.int64)
        np.cumsum(npts, out=offsets[1:])
        U = int(offsets[-1]) - P
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        if seeds.size != U:
            raise ValueError("need one seed per (path, iteration)")
        self._flush_pending()
        _lib.check(self.h, self.lib.pfb_batch_from_lbfgs(self.h, _ptr(seeds)))
        self._P, self._U, self._offsets = P, U, offsets

    def lbfgs_download(self):
        """(offsets [P+1], points [n, T], log_densities [T], gradients [n, T]) of the last lbfgs_batch."""
        npts = self._lb_npts
        offsets = np.zeros(npts.size + 1, dtype=np.int64)
        np.cumsum(npts, out=offsets[1:])
        T = int(offsets[-1])
        X = np.empty((self.n, T), order="F")
        G = np.empty((self.n, T), order="F")
        FX = np.empty(T)
        _lib.check(self.h, self.lib.pfb_lbfgs_download(self.h, _ptr(X), _ptr(G), _ptr(FX)))
        return offsets, X, FX, G

    def lbfgs_ms(self):
        ms = C.c_double(0.0)
        _lib.check(self.h, self.lib.pfb_lbfgs_ms(self.h, C.byref(ms)))
        return ms.value

    def run(self):
        _lib.check(self.h, self.lib.pfb_batch_run(self.h))
        self._raise_cb_error()
        self._poolK = self.K
        self._poolP = None

    def sync(self):
        _lib.check(self.h, self.lib.pfb_batch_sync(self.h))

    def pin(self, *arrays):
        """Page-lock caller-owned output arrays in place (cudaHostRegister) so that reusing them
        through download(into=...) / psis_resample(into=...) gets full-rate D2H copies."""
        for a in arrays:
            if a is not None and a.nbytes > 0:
                rc = self.lib.pfb_host_register(_ptr(a), a.nbytes)
                if rc != 0:
                    raise _lib.PfbError(rc, "cudaHostRegister failed")
        return arrays

    def unpin(self, *arrays):
        for a in arrays:
            if a is not None and a.nbytes > 0:
                self.lib.pfb_host_unregister(_ptr(a))

    def download(self, draws=True, per_draw=False, fit=False, all_draws=False, into=None) -> ElboBatchResult:
        """into: a result of an earlier download of the SAME batch shape whose arrays are reused
        (the caller-owned output buffers of the C ABI; pin() them for full-rate copies)."""
        out, res, succ = self._out_struct(draws, per_draw, fit, all_draws, into)
        _lib.check(self.h, self.lib.pfb_batch_download(self.h, C.byref(out)))
        res.success = succ.astype(bool)
        return res

    def _out_struct(self, draws, per_draw, fit, all_draws, into):
        n, K, P, U, KP = self.n, self.K, self._P, self._U, self.KP
        out = pfb_elbo_out()

        def buf(name, shape, dtype=np.float64, order="C", src=None):
            prev = None if into is None else (src if src is not None else getattr(into, name, None))
            if prev is not None and prev.shape == tuple(np.atleast_1d(shape)) and prev.dtype == dtype and \
                    (prev.flags.f_contiguous if order == "F" else prev.flags.c_contiguous):
                return prev
            return np.empty(shape, dtype=dtype, order=order)

        elbo = buf("elbo", U); se = buf("elbo_se", U)
        best = buf("best_iter", P, np.int64)
        succ = buf("_succ32", P, np.int32)
        rej = buf("n_rejected", P, np.int64)
        out.elbo, out.elbo_se = _ptr(elbo), _ptr(se)
        out.best_iter, out.success, out.n_rejected = _ptr(best), _ptr(succ), _ptr(rej)
        res = ElboBatchResult(self._offsets, elbo, se, best, None, rej)
        res._succ32 = succ
        if draws == "lazy":
            res._lazy_engine = weakref.ref(self)
            self._pending.append(weakref.ref(res))
        elif draws:
            res.draws = buf("draws", (n, K, P), order="F")
            res.draws_logp = buf("draws_logp", (K, P), order="F")
            res.draws_logq = buf("draws_logq", (K, P), order="F")
            out.draws, out.draws_logp, out.draws_logq = _ptr(res.draws), _ptr(res.draws_logp), _ptr(res.draws_logq)
        if per_draw:
            res.logp = buf("logp", (K, U), order="F")
            res.logq = buf("logq", (K, U), order="F")
            out.logp, out.logq = _ptr(res.logp), _ptr(res.logq)
        if fit:
            pf_ = into.fit if (into is not None and into.fit is not None) else {}
            f = dict(mu=buf("mu", (n, P), order="F", src=pf_.get("mu")),
                     alpha=buf("alpha", (n, P), order="F", src=pf_.get("alpha")),
                     vh=buf("vh", (n, KP, P), order="F", src=pf_.get("vh")),
                     T=buf("T", (P, KP, KP), src=pf_.get("T")), Vc=buf("Vc", (P, KP, KP), src=pf_.get("Vc")),
                     logdet=buf("logdet", P, src=pf_.get("logdet")),
                     jeff=buf("jeff", P, np.int32, src=pf_.get("jeff")))
            out.fit_mu, out.fit_alpha, out.fit_vh = _ptr(f["mu"]), _ptr(f["alpha"]), _ptr(f["vh"])
            out.fit_T, out.fit_Vc = _ptr(f["T"]), _ptr(f["Vc"])
            out.fit_logdet, out.fit_jeff = _ptr(f["logdet"]), _ptr(f["jeff"])
            res.fit = f
        if all_draws:
            res.all_draws = np.empty((n, K, U), order="F")
            out.all_draws = _ptr(res.all_draws)
        return out, res, succ

    @staticmethod
    def result_arrays(res):
        """Every host array of a download() result (for pin / unpin)."""
        out = [res.elbo, res.elbo_se, res.best_iter, res.n_rejected, getattr(res, "_succ32", None), res.draws,
               res.draws_logp, res.draws_logq, res.logp, res.logq, res.all_draws]
        if res.fit:
            out += list(res.fit.values())
        return [a for a in out if a is not None]

    def elbo_batch(self, offsets, X, G, seeds, normals=None, draws=True, per_draw=False, fit=False,
                   all_draws=False, into=None) -> ElboBatchResult:
        """Upload, ELBO stage and download as ONE library call (pfb_elbo_batch): the trajectories go up
        in path groups on a copy stream while K1 / K2 already work on the groups that have arrived."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        P = offsets.size - 1
        U = int(offsets[-1]) - P
        X = np.asfortranarray(X, dtype=np.float64)
        G = np.asfortranarray(G, dtype=np.float64)
        if X.shape != (self.n, int(offsets[-1])) or G.shape != X.shape:
            raise ValueError("positions / gradients must be n x offsets[-1]")
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        if seeds.size != U:
            raise ValueError("need one seed per (path, iteration)")
        if normals is not None:
            normals = np.asfortranarray(normals, dtype=np.float64)
            if normals.shape != (self.n, self.K, U):
                raise ValueError("normals must be n x K x U")
        self._flush_pending()
        self._P, self._U, self._offsets = P, U, offsets.copy()
        out, res, succ = self._out_struct(draws, per_draw, fit, all_draws, into)
        _lib.check(self.h, self.lib.pfb_elbo_batch(self.h, self.n, P, _ptr(offsets), _ptr(X), _ptr(G), _ptr(seeds),
                                                   _ptr(normals), C.byref(out)))
        self._raise_cb_error()
        self._poolK = self.K
        self._poolP = None
        res.success = succ.astype(bool)
        return res

    def fit_only(self, best_iter):
        """K1 + K2 with caller-given best iterations (resample() re-entry)."""
        bi = np.ascontiguousarray(best_iter, dtype=np.int64)
        if bi.size != self._P:
            raise ValueError("need one best_iter per path")
        self._flush_pending()  # the device pool is dropped: results that still point at it get their draws now
        _lib.check(self.h, self.lib.pfb_batch_fit_only(self.h, _ptr(bi)))

    def draw_from_fits(self, K_new, seeds, keep_as_pool=False, want_draws=True):
        """K_new fresh draws per path from its best-iteration normal: (draws [n, K_new, P], logp, logq)."""
        sd = np.ascontiguousarray(seeds, dtype=np.uint64)
        if sd.size != self._P:
            raise ValueError("need one seed per path")
        P = self._P
        if keep_as_pool:
            self._flush_pending()
        draws = np.empty((self.n, int(K_new), P), order="F") if want_draws else None
        lp = np.empty((int(K_new), P), order="F")
        lq = np.empty((int(K_new), P), order="F")
        _lib.check(self.h, self.lib.pfb_draw_from_fits(self.h, int(K_new), _ptr(sd), _ptr(draws), _ptr(lp), _ptr(lq),
                                                       int(bool(keep_as_pool))))
        self._raise_cb_error()
        if keep_as_pool:
            self._poolK = int(K_new)
        return draws, lp, lq

    def unit_draws(self, units):
        """(draws [n, K, m], logp [K, m], logq [K, m]) of the given 0-based units of the current batch,
        regenerated on the device from their seeds: ELBOEstimate.draws etc. (src/elbo.jl:22-29) on
        demand.  `ElboBatchResult.unit_slice(p).start + l - 1` is the unit of iteration l of path p."""
        u = np.ascontiguousarray(units, dtype=np.int32)
        m = u.size
        draws = np.empty((self.n, self.K, m), order="F")
        lp = np.empty((self.K, m), order="F")
        lq = np.empty((self.K, m), order="F")
        _lib.check(self.h, self.lib.pfb_unit_draws(self.h, m, _ptr(u), _ptr(draws), _ptr(lp), _ptr(lq)))
        self._raise_cb_error()
        return draws, lp, lq

    def unit_fits(self, units):
        """fit_distributions[l + 1] (src/singlepath.jl:64) of the given 0-based units of the current batch
        in the reference's WoodburyPDMat form: dict(mu [n, m], alpha [n, m], vh [n, KP, m], T [m, KP, KP],
        Vc [m, KP, KP], logdet [m], jeff [m]).  Exported on demand instead of kept for every iteration."""
        u = np.ascontiguousarray(units, dtype=np.int32)
        m, n, KP = u.size, self.n, self.KP
        f = dict(mu=np.empty((n, m), order="F"), alpha=np.empty((n, m), order="F"),
                 vh=np.empty((n, KP, m), order="F"), T=np.empty((m, KP, KP)), Vc=np.empty((m, KP, KP)),
                 logdet=np.empty(m), jeff=np.empty(m, dtype=np.int32))
        _lib.check(self.h, self.lib.pfb_unit_fits(self.h, m, _ptr(u), _ptr(f["mu"]), _ptr(f["alpha"]), _ptr(f["vh"]),
                                                  _ptr(f["T"]), _ptr(f["Vc"]), _ptr(f["logdet"]), _ptr(f["jeff"])))
        return f

    # ---- multi-GPU (one engine per GPU; NCCL behind the C ABI) ---------------------------------
    def comm_init(self, group=None):
        """Join the engines of a torch.distributed process group into one NCCL communicator owned by the
        library (pfb_comm_init): rank 0 creates the id, the group broadcasts its 128 bytes."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            rc = self.lib.pfb_comm_unique_id(_ptr(buf))
            if rc != 0:
                raise _lib.PfbError(rc, (self.lib.pfb_last_error(None) or b"").decode())
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.from_numpy(buf).to(dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        buf = t.cpu().numpy()
        _lib.check(self.h, self.lib.pfb_comm_init(self.h, _ptr(buf), rank, world))
        self._comm = (rank, world)

    def pool_set(self, P, K_run, draws, logp, logq):
        """Install a device pool assembled on the host (pfb_pool_set): draws [n, K_run, P] or None."""
        d = None if draws is None else np.asfortranarray(draws, dtype=np.float64)
        lp = np.asfortranarray(logp, dtype=np.float64)
        lq = np.asfortranarray(logq, dtype=np.float64)
        self._flush_pending()
        _lib.check(self.h, self.lib.pfb_pool_set(self.h, int(P), int(K_run), _ptr(d), _ptr(lp), _ptr(lq)))
        self._poolK, self._poolP = int(K_run), int(P)

    def pool_exchange_resample(self, paths_per_rank, seed, ndraws, importance=True, replace=True, want_weights=True,
                               into=None):
        """PSIS + resampling over the pools of ALL ranks (pfb_pool_exchange_resample): all-gather of the
        per-draw log densities, replicated PSIS / index draw, owned columns regenerated, sum-reduce.
        Every rank gets the same dict as psis_resample()."""
        ppr = np.ascontiguousarray(paths_per_rank, dtype=np.int32)
        N = int(ppr.sum()) * self._poolK
        out, r = self._resample_out(N, ndraws, importance, True, want_weights, into=into)
        _lib.check(self.h, self.lib.pfb_pool_exchange_resample(
            self.h, _ptr(ppr), C.c_uint64(int(seed)), int(ndraws), int(bool(importance)), int(bool(replace)),
            C.byref(out)))
        return self._finish(r)

    def timings(self):
        ms = np.zeros(6)
        nl = self.lib.pfb_get_timings(self.h, _ptr(ms))
        return dict(k1=ms[0], k2=ms[1], k3=ms[2], k4=ms[3], k5=ms[4], total=ms[5], launches=nl)

    def pool_materialize(self):
        """Materialise the pool's draws on the device (needed before device_view().pool_draws is read)."""
        _lib.check(self.h, self.lib.pfb_pool_materialize(self.h))

    def pool_columns_device(self, m, d_inds, base, d_out):
        """Regenerate this engine's share of the globally indexed (1-based) pool columns d_inds[m]
        (device int64 pointer) into d_out [n x m] (device pointer); see pfb_pool_columns_device."""
        _lib.check(self.h, self.lib.pfb_pool_columns_device(self.h, int(m), C.c_void_p(int(d_inds)), int(base),
                                                            C.c_void_p(int(d_out))))

    def device_view(self) -> pfb_device_view:
        v = pfb_device_view()
        _lib.check(self.h, self.lib.pfb_batch_device_view(self.h, C.byref(v)))
        return v

    # ---- PSIS + resample stage -------------------------------------------------------------
    def _resample_out(self, N, ndraws, importance, want_draws, want_weights=True, into=None):
        out = pfb_resample_out()
        if into is not None and into["inds"].shape == (ndraws,) and \
                (not (importance and want_weights) or into.get("weights", np.empty(0)).shape == (N,)) and \
                (not want_draws or into.get("draws", np.empty((0, 0))).shape == (self.n, ndraws)):
            r = dict(into)  # reuse the caller's (possibly pinned) output arrays
            r["pareto_k"] = np.full(1, np.nan); r["tail_len"] = np.zeros(1, dtype=np.int64)
            out.inds, out.ids = _ptr(r["inds"]), _ptr(r["ids"])
            out.pareto_k, out.tail_len = _ptr(r["pareto_k"]), _ptr(r["tail_len"])
            if importance and want_weights:
                out.log_weights, out.weights = _ptr(r["log_weights"]), _ptr(r["weights"])
            if want_draws:
                out.draws = _ptr(r["draws"])
            return out, r
        r = dict(inds=np.empty(ndraws, dtype=np.int64), ids=np.empty(ndraws, dtype=np.int64),
                 pareto_k=np.full(1, np.nan), tail_len=np.zeros(1, dtype=np.int64))
        out.inds, out.ids = _ptr(r["inds"]), _ptr(r["ids"])
        out.pareto_k, out.tail_len = _ptr(r["pareto_k"]), _ptr(r["tail_len"])
        if importance and want_weights:
            r["log_weights"] = np.empty(N)
            r["weights"] = np.empty(N)
            out.log_weights, out.weights = _ptr(r["log_weights"]), _ptr(r["weights"])
        if want_draws:
            r["draws"] = np.empty((self.n, ndraws), order="F")
            out.draws = _ptr(r["draws"])
        return out, r

    @staticmethod
    def _finish(r):
        r["pareto_k"] = float(r["pareto_k"][0])
        r["tail_len"] = int(r["tail_len"][0])
        return r

    def psis_resample(self, seed, ndraws, importance=True, replace=True, into=None):
        """On the pool of the last batch / the last draw_from_fits(keep_as_pool=True) (device resident).
        into: the dict of an earlier call with the same shapes, whose arrays are reused."""
        N = (self._poolP if self._poolP is not None else self._P) * self._poolK
        out, r = self._resample_out(N, ndraws, importance, True, into=into)
        _lib.check(self.h, self.lib.pfb_psis_resample(self.h, C.c_uint64(int(seed)), int(ndraws),
                                                      int(bool(importance)), int(bool(replace)), C.byref(out)))
        return self._finish(r)

    def psis_resample_host(self, log_ratios, K_run, seed, ndraws, importance=True, pool=None, N=None, replace=True):
        lr = None if log_ratios is None else np.ascontiguousarray(log_ratios, dtype=np.float64)
        if pool is not None:
            pool = np.asfortranarray(pool, dtype=np.float64)
            N = pool.shape[1]
        elif lr is not None:
            N = lr.size
        elif N is None:
            raise ValueError("pool size N is needed when neither log_ratios nor pool is given")
        out, r = self._resample_out(N, ndraws, importance, pool is not None)
        _lib.check(self.h, self.lib.pfb_psis_resample_host(
            self.h, self.n, N, int(K_run), _ptr(lr), _ptr(pool), C.c_uint64(int(seed)), int(ndraws),
            int(bool(importance)), int(bool(replace)), C.byref(out)))
        return self._finish(r)

    def psis_resample_device(self, N, K_run, d_logp, d_logq, d_pool, seed, ndraws, importance=True, replace=True,
                             want_weights=True):
        """d_* are raw device pointers (ints), e.g. torch tensors' data_ptr().  want_weights=False
        leaves the N-sized weight vectors on the device (only indices, ids, k-hat come back)."""
        out, r = self._resample_out(N, ndraws, importance, d_pool is not None, want_weights)
        _lib.check(self.h, self.lib.pfb_psis_resample_device(
            self.h, self.n, int(N), int(K_run), C.c_void_p(d_logp), C.c_void_p(d_logq),
            C.c_void_p(d_pool) if d_pool else None, C.c_uint64(int(seed)), int(ndraws),
            int(bool(importance)), int(bool(replace)), C.byref(out)))
        return self._finish(r)
