#!/usr/bin/env python
"""Coefficients of the polynomial R(z) ~ (asin(sqrt z) / sqrt z - 1) / z on z in [0, 1/4] used by the device-side
acos / asin of the penumbra evaluation (leo_core.cuh: asin_poly): Chebyshev interpolation in 60-digit arithmetic,
converted to the monomial basis; prints the max error of s + s z R(z) against asin(s) for s in [0, 1/2]."""
import sys
import mpmath as mp
mp.mp.dps = 60
N = int(sys.argv[1]) if len(sys.argv) > 1 else 15
a, b = mp.mpf(0), mp.mpf(1) / 4


def f(z):
    if z < mp.mpf(10) ** -40:
        return mp.mpf(1) / 6 + 3 * z / 40
    s = mp.sqrt(z)
    return (mp.asin(s) / s - 1) / z


# Chebyshev nodes / coefficients on [a, b]
n = N + 1
xs = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
fs = [f((b - a) / 2 * x + (a + b) / 2) for x in xs]
c = [2 / mp.mpf(n) * sum(fs[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) for j in range(n)]
c[0] /= 2
# Chebyshev T_j(t) as monomials in t, then t = (2 z - (a + b)) / (b - a) = 8 z - 1
T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
for j in range(2, n):
    prev, pp = T[j - 1], T[j - 2]
    cur = [mp.mpf(0)] + [2 * v for v in prev]
    for i, v in enumerate(pp):
        cur[i] -= v
    T.append(cur)
pt = [mp.mpf(0)] * n
for j in range(n):
    for i, v in enumerate(T[j]):
        pt[i] += c[j] * v
# substitute t = 8 z - 1
pz = [mp.mpf(0)] * n
for i, v in enumerate(pt):
    # (8z - 1)^i
    for k in range(i + 1):
        pz[k] += v * mp.binomial(i, k) * (mp.mpf(8) ** k) * ((-1) ** (i - k))
coef = [float(v) for v in pz]


def R(z):
    r = 0.0
    for v in reversed(coef):
        r = r * z + v
    return r


worst = 0
for k in range(0, 2001):
    s = 0.5 * k / 2000
    z = s * s
    got = s + s * z * R(z)
    worst = max(worst, abs(mp.mpf(got) - mp.asin(mp.mpf(s))))
print("degree", N, "max abs error of asin on [0, 0.5]:", mp.nstr(worst, 3))
print(", ".join(repr(v) for v in coef))
