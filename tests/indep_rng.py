"""A second, separately written statement of the engine's normal-variate contract (RNG contract v2), in
pure Python: Philox4x32-7 on Python integers, the 1024-layer ziggurat on the tables recomputed by the
generator script (scripts/gen_zig_tables.py), exact rational arithmetic for the one rounding of the fast
path, libm for the slow path.  The contract's own source (pathfinder_b200/csrc/pf_rng.h) is compiled into
both the kernels and the oracle; this file shares no code with it, so agreement here means the C source
says what DESIGN.md section 2 says:

  element (row i, draw k) of seed s:  call Philox4x32-7 with the fixed key on the counter
      (i >> 1 | stream << 28,  (k >> 4) * 8 + (k & 7),  s_lo + call,  s_hi),   stream = 0, call = 0;
  its word 2 * ((k >> 3) & 1) + (i & 1) is the variate's 32 bits: bit 31 sign, bits 21-30 layer,
  bits 0-19 mantissa j;  x = j 2^-20 x_layer;  accepted at once when j < kq_layer (99.57 %);
  otherwise the element continues on its private stream 1 + word (call = 0, 1, ...): layer 0 -> Marsaglia's
  exponential tail beyond r, else the wedge test y < exp(-x^2 / 2) with y uniform between f(x_layer) and
  f(x_layer + 1); a rejected candidate restarts from word 2 of the same call.
"""
import math
import struct
from fractions import Fraction

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
KEY = (0xA4093822, 0x299F31D0)
MASK = 0xFFFFFFFF


def philox4x32(rounds, ctr, key=KEY):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(rounds):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = (p1 >> 32) ^ c1 ^ k0, p1 & MASK, (p0 >> 32) ^ c3 ^ k1, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def _dbl(u):
    return struct.unpack("<d", struct.pack("<Q", u))[0]


def _tables():
    import importlib.util
    import pathlib

    path = pathlib.Path(__file__).resolve().parents[1] / "scripts" / "gen_zig_tables.py"
    spec = importlib.util.spec_from_file_location("gen_zig_tables", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _, _, _, _, xk, xp, fx = mod.tables()
    return xk, xp, fx


_T = None


def _tables_cached():
    global _T
    if _T is None:
        _T = _tables()
    return _T


def _call(rp, stream, dp, seed, call):
    return philox4x32(7, (rp | stream << 28, dp, ((seed & MASK) + call) & MASK, seed >> 32))


def _u01(bits):
    return ((bits >> 11) + 0.5) * 2.0 ** -53


def _value(w, edge):
    j = w & 0xFFFFF
    x = float(Fraction(j, 1 << 20) * Fraction(edge))  # j 2^-20 x_layer, rounded once
    return -x if w >> 31 else x


def normal_elem(seed, i, k):
    """(variate, took_the_slow_path)."""
    xk, xp, fx = _tables_cached()
    rp, dp, word = i >> 1, (k >> 4) * 8 + (k & 7), 2 * ((k >> 3) & 1) + (i & 1)
    w = _call(rp, 0, dp, seed, 0)[word]
    layer = (w >> 21) & 1023
    if (w & 0xFFFFF) < (xk[layer] & 0xFFFFF):
        return _value(w, xp[layer]), False
    stream, call = 1 + word, 0
    while True:
        layer = (w >> 21) & 1023
        z = _value(w, xp[layer])
        if (w & 0xFFFFF) < (xk[layer] & 0xFFFFF):
            return z, True
        neg = w >> 31
        if layer == 0:
            r = xp[1]
            while True:
                o = _call(rp, stream, dp, seed, call)
                call += 1
                xt = -math.log(_u01(o[0] | o[1] << 32)) / r
                yt = -math.log(_u01(o[2] | o[3] << 32))
                if yt + yt > xt * xt:
                    return (-(r + xt) if neg else r + xt), True
        o = _call(rp, stream, dp, seed, call)
        call += 1
        y = fx[layer] + _u01(o[0] | o[1] << 32) * (fx[layer + 1] - fx[layer])
        if y < math.exp(-0.5 * z * z):
            return z, True
        w = o[2]
