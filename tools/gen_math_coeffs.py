"""Derives the polynomial coefficients used by microclimf_b200/csrc/mcf_math.cuh with mpmath
(Chebyshev-node interpolation at 60 digits ~ minimax to within a small factor) and checks the
resulting double-precision algorithms against mpmath on random points.

  exp : exp(r) on |r| <= ln2/2, degree 11
  log : log(1+f) = f - s*(f - z*P(z)), s = f/(2+f), z = s^2, P(z) ~ 2/3 + 2z/5 + ... on z in [0, 0.0295], degree 6
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_fit(fn, a, b, deg):
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = fn(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def hexd(v):
    return float(v).hex()


ln2 = mp.log(2)
ce = cheb_fit(mp.exp, -ln2 / 2, ln2 / 2, 11)
print("// exp(r), |r| <= ln2/2, degree 11")
for i, c in enumerate(ce):
    print(f"    {float(c)!r},  // c{i}")


def P(z):
    if z == 0:
        return mp.mpf(2) / 3
    s = mp.sqrt(z)
    return (2 * mp.atanh(s) - 2 * s) / (s * z)


smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
cl = cheb_fit(P, mp.mpf(0), smax ** 2 * mp.mpf("1.02"), 6)
print("// log: P(z), degree 6")
for i, c in enumerate(cl):
    print(f"    {float(c)!r},  // L{i}")
hi = float(ln2)
hi = np.float64(hi).view(np.uint64) & np.uint64(0xFFFFFFFFF8000000)
hi = hi.view(np.float64)
lo = float(ln2 - mp.mpf(float(hi)))
print("ln2_hi", repr(float(hi)), "ln2_lo", repr(lo), "log2e", repr(float(1 / ln2)))

# ---- numpy emulation of the algorithms (no fma: errors here are upper bounds of the fma version)
cef = np.array([float(c) for c in ce])
clf = np.array([float(c) for c in cl])


def fexp(x):
    k = np.rint(x * float(1 / ln2))
    r = (x - k * float(hi)) - k * lo
    p = np.zeros_like(r)
    for c in cef[::-1]:
        p = p * r + c
    return np.ldexp(p, k.astype(np.int64))


def flog(x):
    m, e = np.frexp(x)  # m in [0.5, 1)
    big = m < np.sqrt(0.5)
    m = np.where(big, 2 * m, m)
    e = np.where(big, e - 1, e)
    f = m - 1
    s = f / (2 + f)
    z = s * s
    p = np.zeros_like(z)
    for c in clf[::-1]:
        p = p * z + c
    R = z * p
    return e * float(hi) + ((f - s * (f - R)) + e * lo)


rng = np.random.default_rng(0)
x = rng.uniform(-700, 700, 200000)
ref = np.array([float(mp.exp(mp.mpf(v))) for v in x[:20000]])
print("exp max rel err", np.max(np.abs(fexp(x[:20000]) / ref - 1)))
x = np.exp(rng.uniform(-50, 50, 20000))
ref = np.array([float(mp.log(mp.mpf(v))) for v in x])
print("log max abs/rel err", np.max(np.abs(flog(x) - ref) / np.maximum(np.abs(ref), 1e-300)))
x = rng.uniform(0.5, 2.0, 20000)
ref = np.array([float(mp.log(mp.mpf(v))) for v in x])
print("log near 1 max rel err", np.max(np.abs(flog(x) - ref) / np.abs(ref)))
