"""Derives the polynomial of mlb_math.cuh's asinSmall: asin(x) = x + x z P(z), z = x^2, for |x| <= 0.5, P of degree DEG fitted at
Chebyshev nodes in 60-digit arithmetic (mpmath) and rounded to double; prints the coefficients and the worst error in ulps of the
double-precision Horner evaluation against the 60-digit value over a dense sample. CPU only, run once:  python tools/fit_asin.py"""
import sys
import mpmath as mp
import numpy as np
mp.mp.dps = 60
DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 15
ZMAX = mp.mpf(sys.argv[2]) if len(sys.argv) > 2 else mp.mpf("0.25")

def f(z):
    if z == 0:
        return mp.mpf(1) / 6
    x = mp.sqrt(z)
    return (mp.asin(x) / x - 1) / z

n = DEG + 1
nodes = [ZMAX / 2 * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * n))) for k in range(n)]
A = mp.matrix(n, n)
b = mp.matrix(n, 1)
for i, z in enumerate(nodes):
    for j in range(n):
        A[i, j] = z ** j
    b[i] = f(z)
c = mp.lu_solve(A, b)
coef = [float(c[j]) for j in range(n)]
print("static __constant__ double kAsinP[%d] = {" % n)
print("    " + ", ".join(float.hex(v) for v in coef))
print("};   /* decimal: " + ", ".join(repr(v) for v in coef) + " */")

# error of the double evaluation (Horner without FMA here; the device uses fma: not worse)
rng = np.random.default_rng(1)
xs = np.concatenate([rng.uniform(-0.5, 0.5, 200000), np.linspace(-0.5, 0.5, 20001), 10.0 ** rng.uniform(-12, -1, 20000)])
worst = 0.0
for x in xs[::7]:
    z = x * x
    p = coef[-1]
    for v in coef[-2::-1]:
        p = p * z + v
    y = x + x * (z * p)
    exact = mp.asin(mp.mpf(float(x)))
    ulp = np.spacing(abs(float(exact))) if exact != 0 else 1
    worst = max(worst, abs(float((mp.mpf(float(y)) - exact) / ulp)))
polyErr = max(abs(float((sum(mp.mpf(cv) * z ** j for j, cv in enumerate(coef)) - f(z)) / f(z))) for z in [ZMAX * k / 400 for k in range(401)])
print("degree", DEG, "max relative error of P:", polyErr, " worst error of the double evaluation: %.3f ulp" % worst)
