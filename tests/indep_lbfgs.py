"""A second, separately written statement of the engine's L-BFGS trajectory contract, in plain NumPy.

The contract's own source (pathfinder_b200/csrc/pf_lbfgs.h) is compiled by nvcc into K0 and by g++ into
the oracle, so "K0 == oracle, bit for bit" compares one source with itself under two compilers.  This
file shares no code with it: it is the algorithm as DESIGN.md section 2 describes it — two-loop recursion
with H0 = (s'y / y'y) I over a ring of `history_length` pairs, a strong-Wolfe bracketing / zoom line search
with safeguarded cubic interpolation (c1 = 1e-4, c2 = 0.9, first trial 1/|g| on a restart, x4 expansion,
at most 40 evaluations), the trace and stopping rules of src/optimize.jl:94-105 — written with NumPy
reductions (np.dot: another summation order, no emulated thread layout, libm's exp).  Agreement is
therefore up to rounding, which the tests bound.
"""
import numpy as np

GTOL, FTOL, C1, C2, MAXLS = 1e-8, 1e-14, 1e-4, 0.9, 40


def density(kind, **kw):
    """(log p, gradient of log p) closures of the registered closed-form families."""
    if kind == "isonormal":
        return lambda x: (-0.5 * np.dot(x, x), -x)
    if kind == "diagnormal":
        mean, sd = kw["mean"], kw["sd"]
        c0 = -np.sum(np.log(sd)) - 0.5 * mean.size * np.log(2 * np.pi)

        def f(x):
            z = (x - mean) / sd
            return -0.5 * np.dot(z, z) + c0, -z / sd
        return f
    if kind == "funnel":  # docs/src/examples/quickstart.md:229-234
        def f(x):
            n, x0 = x.size, x[0]
            e, ss = np.exp(-x0), np.dot(x[1:], x[1:])
            lp = -0.5 * ((x0 / 3.0) ** 2 + (n - 1) * x0 + e * ss)
            g = np.empty(n)
            g[0] = -0.5 * (2.0 * x0 / 9.0 + (n - 1) - e * ss)
            g[1:] = -e * x[1:]
            return lp, g
        return f
    if kind == "densenormal":
        mean, prec = kw["mean"], kw["prec"]

        def f(x):
            z = x - mean
            pz = prec @ z
            return -0.5 * np.dot(z, pz), -pz
        return f
    raise ValueError(kind)


def lbfgs_path(fun, x0, history_length=6, maxiters=1000, max_points=None, gtol=GTOL, ftol=FTOL):
    """Returns (points [n, L+1], log densities [L+1], gradients of log p [n, L+1], status, evaluations)."""
    max_points = min(maxiters + 1, max_points or maxiters + 1)
    x = np.array(x0, dtype=float)
    lp, g = fun(x)
    X, FX, G, nev = [x.copy()], [lp], [g.copy()], 1
    finite = lambda v, gr: np.isfinite(v) and np.all(np.isfinite(gr))
    if not finite(lp, g):
        return np.stack(X, 1), np.array(FX), np.stack(G, 1), "nonfinite", nev
    if np.max(np.abs(g)) <= gtol:
        return np.stack(X, 1), np.array(FX), np.stack(G, 1), "gtol", nev
    pairs = []  # (s, y, rho), oldest first
    gamma, status, f = 1.0, "maxiters", -lp
    for _ in range(maxiters):
        if len(X) >= max_points:
            break
        grad_f = -g  # the minimised function is -log p
        # two-loop recursion
        q = grad_f.copy()
        alphas = []
        for s, y, rho in reversed(pairs):
            a = rho * np.dot(s, q)
            alphas.append(a)
            q -= a * y
        if pairs:
            q *= gamma
        for (s, y, rho), a in zip(pairs, reversed(alphas)):
            b = rho * np.dot(y, q)
            q += (a - b) * s
        d = -q
        dg = np.dot(grad_f, d)
        if not dg < 0.0:
            pairs = []
            d = -grad_f
            dg = -np.dot(grad_f, grad_f)
            if not dg < 0.0:
                status = "linesearch"
                break
        a_init = 1.0
        if not pairs:
            gn = np.sqrt(np.dot(grad_f, grad_f))
            a_init = 1.0 / gn if gn > 1.0 else 1.0

        cache = {}

        def phi_at(a):
            nonlocal nev
            xt = x + a * d
            lpt, gt = fun(xt)
            nev += 1
            ph = -lpt if np.isfinite(lpt) else np.inf
            cache["last"] = (a, xt, lpt, gt)
            return ph, np.dot(-gt, d)

        phi0, dphi0 = f, dg
        lo = (0.0, phi0, dphi0)
        hi = (0.0, phi0, dphi0)
        acc, bracket, ls = None, False, 0
        a, prev = a_init, (0.0, phi0, dphi0)
        while ls < MAXLS:
            ph, dph = phi_at(a)
            ls += 1
            if ph > phi0 + C1 * a * dphi0 or (ls > 1 and ph >= prev[1]):
                lo, hi, bracket = prev, (a, ph, dph), True
                break
            if abs(dph) <= -C2 * dphi0:
                acc = a
                break
            if dph >= 0.0:
                lo, hi, bracket = (a, ph, dph), prev, True
                break
            prev = (a, ph, dph)
            lo = prev
            a *= 4.0
        while bracket and acc is None and ls < MAXLS:
            (alo, plo, dlo), (ahi, phi_h, dhi) = lo, hi
            w = ahi - alo
            with np.errstate(all="ignore"):
                d1 = dlo + dhi - 3.0 * (plo - phi_h) / (alo - ahi)
                rad = d1 * d1 - dlo * dhi
                d2 = np.sqrt(rad) if w > 0.0 else -np.sqrt(rad)
                at = ahi - w * ((dhi + d2 - d1) / (dhi - dlo + 2.0 * d2))
            lob, hib = min(alo, ahi), max(alo, ahi)
            margin = 0.1 * (hib - lob)
            if not (lob + margin <= at <= hib - margin):
                at = alo + 0.5 * w
            if at == alo or at == ahi:
                break
            ph, dph = phi_at(at)
            ls += 1
            if ph > phi0 + C1 * at * dphi0 or ph >= plo:
                hi = (at, ph, dph)
            else:
                if abs(dph) <= -C2 * dphi0:
                    acc = at
                    break
                if dph * (ahi - alo) >= 0.0:
                    hi = lo
                lo = (at, ph, dph)
        if acc is None:
            if lo[0] > 0.0 and lo[1] <= phi0 + C1 * lo[0] * dphi0 and lo[1] < phi0:
                acc = lo[0]
            else:
                status = "linesearch"
                break
        if cache["last"][0] != acc:
            phi_at(acc)
        _, xt, lpt, gt = cache["last"]
        X.append(xt.copy()); FX.append(lpt); G.append(gt.copy())
        f_new = -lpt if np.isfinite(lpt) else np.inf
        if not np.all(np.isfinite(gt)):
            status = "nonfinite"
            break
        s, y = xt - x, g - gt
        sy, yy = np.dot(s, y), np.dot(y, y)
        if sy > 0.0 and yy > 0.0 and np.isfinite(sy) and np.isfinite(yy):
            pairs.append((s, y, 1.0 / sy))
            pairs = pairs[-history_length:]
            gamma = sy / yy
        x, g = xt, gt
        if np.max(np.abs(gt)) <= gtol:
            status = "gtol"
            break
        if f - f_new <= ftol * max(abs(f), abs(f_new), 1.0):
            status = "ftol"
            break
        f = f_new
    return np.stack(X, 1), np.array(FX), np.stack(G, 1), status, nev
