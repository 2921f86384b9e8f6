"""Near-minimax coefficients for the arcsine kernel of rv2coe.cu:
asin(t) = t + t^3 P(t^2) on t^2 <= ZMAX, P of degree N - 1, fitted at Chebyshev nodes in
60-digit arithmetic and checked in IEEE double with the kernel's Horner order.
usage: python scripts/fit_asin.py [ZMAX=0.25] [N=16]"""
import sys
import mpmath as mp
import numpy as np

mp.mp.dps = 60
zmax = mp.mpf(sys.argv[1]) if len(sys.argv) > 1 else mp.mpf("0.25")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16


def f(z):
    if z == 0:
        return mp.mpf(1) / 6
    t = mp.sqrt(z)
    return (mp.asin(t) - t) / (z * t)


nodes = [zmax / 2 * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * N))) for k in range(N)]
A = mp.matrix(N, N)
b = mp.matrix(N, 1)
for i, z in enumerate(nodes):
    for j in range(N):
        A[i, j] = z ** j
    b[i] = f(z)
c = mp.lu_solve(A, b)
coef = [float(c[j]) for j in range(N)]
print("coefficients (z^0 .. z^%d):" % (N - 1))
for v in coef:
    print("    %.20e," % v)
tmax = float(mp.sqrt(zmax))
t = np.concatenate([np.linspace(-tmax, tmax, 200001), np.logspace(-12, -1, 2000)])
z = t * t
p = np.full_like(t, coef[-1])
for v in coef[-2::-1]:
    p = p * z + v
got = t + (t * z) * p
want = np.array([float(mp.asin(mp.mpf(float(x)))) for x in t[::97]])
err = np.abs(got[::97] - want)
print("max abs error %.3e, max ulp error %.2f" % (err.max(), (err / np.spacing(np.abs(want) + 1e-300)).max()))
