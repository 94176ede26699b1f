"""Tables and accuracy check for the table-driven exp / log of microclimf_b200/csrc/mcf_math.cuh.

  exp(x) = 2^(k/32) * exp(r),  k = round(32 x / ln2),  r = x - k ln2/32, |r| <= ln2/64:  T[k & 31] * P5(r) * 2^(k >> 5)
  log(x) = e ln2 + log(c_j) + log1p(r),  m = mantissa in [1, 2), j = top 6 mantissa bits, r = m / c_j - 1, |r| < 2^-7

Emulates the double-precision evaluation order in numpy (FMAs emulated in long double where it matters) and
compares with mpmath.  Prints the tables as C initialisers.
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 50
ln2 = mp.log(2)

T = [float(mp.power(2, mp.mpf(j) / 32)) for j in range(32)]
invc = [float(1 / (1 + (mp.mpf(j) + mp.mpf(1) / 2) / 64)) for j in range(64)]
logc = [float(-mp.log(mp.mpf(ic))) for ic in invc]  # log of the ROUNDED reciprocal's inverse

C32 = float(32 / ln2)
L32 = float(ln2 / 32)
LN2 = float(ln2)
MAGIC = 6755399441055744.0


def fma(a, b, c):
    return np.asarray(np.longdouble(a) * np.longdouble(b) + np.longdouble(c), dtype=np.float64)


def mexp(x):
    x = np.asarray(x, dtype=np.float64)
    t = fma(x, C32, MAGIC)
    k = (t.view(np.int64) & 0xFFFFFFFF).astype(np.int64)
    k = np.where(k >= 2 ** 31, k - 2 ** 32, k)
    kf = t - MAGIC
    r = fma(kf, -L32, x)
    r2 = r * r
    a = fma(r, 1 / 6, 0.5)
    b = fma(r, 1 / 120, 1 / 24)
    c = fma(r2, b, a)
    d = r + 1.0
    p = fma(r2, c, d)
    tab = np.array(T)[k & 31]
    return np.ldexp(tab * p, (k >> 5).astype(np.int64))


def mlog(x):
    x = np.asarray(x, dtype=np.float64)
    bits = x.view(np.int64)
    hi = (bits >> 32).astype(np.int64)
    e = (hi >> 20) - 1023
    j = (hi >> 14) & 63
    mbits = (bits & 0x000FFFFFFFFFFFFF) | (0x3FF << 52)
    m = mbits.view(np.float64)
    ic = np.array(invc)[j]
    lc = np.array(logc)[j]
    r = fma(m, ic, -1.0)
    r2 = r * r
    a = fma(r, 1 / 3, -0.5)
    b = fma(r, 1 / 5, -0.25)
    c = fma(r2, -1 / 6, b)
    q = fma(r2, c, a)
    res = fma(r2, q, r)
    return fma(e.astype(np.float64), LN2, lc + res)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-20, 20, 20000), rng.uniform(-1, 1, 5000)])
    got = mexp(xs)
    err = max(abs((mp.mpf(float(g)) - mp.exp(mp.mpf(float(x)))) / mp.exp(mp.mpf(float(x)))) for g, x in zip(got[::7], xs[::7]))
    print("// exp: max relative error", float(err))
    xs = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(1e-6, 10, 20000), rng.uniform(0.9, 1.1, 5000)])
    got = mlog(xs)
    ea = max(abs(mp.mpf(float(g)) - mp.log(mp.mpf(float(x)))) / max(1, abs(mp.log(mp.mpf(float(x))))) for g, x in zip(got[::7], xs[::7]))
    print("// log: max error relative to max(1, |log x|)", float(ea))
    print("constexpr double kExp32C = %r, kLn2_32 = %r;" % (C32, L32))
    print("__device__ const double kExpTab[32] = {")
    for j in range(0, 32, 4):
        print("    " + ", ".join(repr(v) for v in T[j:j + 4]) + ",")
    print("};")
    print("__device__ const double2 kLogTab[64] = { // {1 / c_j, log c_j}, c_j = 1 + (j + 0.5) / 64")
    for j in range(0, 64, 2):
        print("    " + ", ".join("{%r, %r}" % (invc[k], logc[k]) for k in range(j, j + 2)) + ",")
    print("};")
