"""The G-step's fixed-point summation written out in numpy / Python integers — the specification the CUDA kernels are
compared with bit for bit (tests/test_gpu_parity.py) and the stand-in the world-size-2 gloo test drives through the sharded host loop
(tests/test_sharding_gloo.py).  Test infrastructure.

  scale_words(E, t)                 -> [bits(max finite |E|), bits(max finite |t|), 1 if a t is not finite else 0, 0]   (ranks: elementwise MAX)
  accumulate(data, t, E, words)     -> limbs int64[768], special float64[256], counts int64[256]                         (ranks: SUM)
  finish(words, limbs, special, n)  -> G float64[256]  (sum / count, gaps extrapolated as in main_responseCalib.cpp:300-304)
"""
import math

import numpy as np

LIMIT = float(2 ** 48)          # scaled products stay below it
LIMB_BITS = 43


def _bits(x):
    return int(np.float64(x).view(np.int64))


def _from_bits(b):
    return float(np.int64(b).view(np.float64))


def scale_words(E, t):
    e = np.abs(E[np.isfinite(E)])
    tt = np.abs(t[np.isfinite(t)])
    return np.array([_bits(e.max()) if e.size else 0, _bits(tt.max()) if tt.size else 0, int(not np.all(np.isfinite(t))), 0], np.int64)


def exponent(words):
    """s with |E*t| * 2^s < 2^48 for every finite sample; None if there is no such bound."""
    p = _from_bits(words[0]) * _from_bits(words[1])
    if not math.isfinite(p) or words[2]:
        return None
    if not p > 0.0:
        return 0
    s = 48 - math.frexp(p)[1]          # frexp: p = m * 2^e with 0.5 <= m < 1, i.e. e = ilogb(p) + 1
    return max(-1000, min(1000, s))


def accumulate(data, t, E, words):
    s = exponent(words)
    scale = 1.0 if s is None else math.ldexp(1.0, s)
    with np.errstate(invalid="ignore", over="ignore"):
        x = (E * scale)[None, :] * t[:, None]          # (E * 2^s) * t, rounded like the reference's E * t
        keep = data != 255                             # :293
        ok = keep & (np.abs(x) < LIMIT)                # False for NaN
        counts = np.bincount(data[keep].ravel(), minlength=256)[:256].astype(np.int64)
        special = np.zeros(256)
        odd = keep & ~ok
        np.add.at(special, data[odd], x[odd] / scale)
    totals = [0] * 256
    v = np.rint(x[ok]).astype(np.int64)                # round to nearest even, exact below 2^48
    for b, val in zip(data[ok].tolist(), v.tolist()):
        totals[b] += val
    limbs = np.zeros(768, np.int64)
    mask = (1 << LIMB_BITS) - 1
    for b in range(256):
        limbs[b], limbs[256 + b], limbs[512 + b] = totals[b] & mask, (totals[b] >> LIMB_BITS) & mask, totals[b] >> (2 * LIMB_BITS)
    limbs[255] = limbs[511] = limbs[767] = 0           # value 255 is never summed
    return limbs, special, counts


def sums(words, limbs, special):
    s = exponent(words) or 0
    out = np.zeros(256)
    for b in range(256):
        v = int(limbs[b]) + (int(limbs[256 + b]) << LIMB_BITS) + (int(limbs[512 + b]) << (2 * LIMB_BITS))
        lo, hi = v & ((1 << 64) - 1), v >> 64
        out[b] = math.ldexp(float(hi) * 18446744073709551616.0 + float(lo), -s) + special[b]
    return out


def finish(words, limbs, special, counts):
    with np.errstate(divide="ignore", invalid="ignore"):
        g = sums(words, limbs, special) / counts.astype(np.float64)
    for i in range(2, 256):                             # main_responseCalib.cpp:300-304, sequential
        if not np.isfinite(g[i]):
            g[i] = g[i - 1] + (g[i - 1] - g[i - 2])
    return g


def gstep(data, t, E):
    """The whole G-step on one rank."""
    w = scale_words(E, t)
    limbs, special, counts = accumulate(data, t, E, w)
    return finish(w, limbs, special, counts)
