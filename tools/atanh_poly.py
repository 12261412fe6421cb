"""Coefficients of the log1p kernel in gadget3_b200/csrc/g3_math.cuh: log1p(e) = t + t z q(z), s = e / (2 + e), t = 2 s, z = s^2,
q(z) = (atanh(sqrt z) / sqrt z - 1) / z on [0, 1/9], by polynomial interpolation at Chebyshev nodes (mpmath, 60 digits).
Prints the coefficients (highest degree first, C99 hex floats) and the error of the double-rounded polynomial."""
import sys
import mpmath as mp
mp.mp.dps = 60
n = int(sys.argv[1]) if len(sys.argv) > 1 else 9


def f(z):
    z = mp.mpf(z)
    if z < mp.mpf(10) ** -40:
        return mp.mpf(1) / 3 + z / 5
    s = mp.sqrt(z)
    return (mp.atanh(s) / s - 1) / z


a, b = mp.mpf(0), mp.mpf(1) / 9
xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * i + 1) / (2 * (n + 1))) for i in range(n + 1)]
V = mp.matrix(n + 1, n + 1)
for i, x in enumerate(xs):
    for j in range(n + 1):
        V[i, j] = x ** j
c = mp.lu_solve(V, mp.matrix([f(x) for x in xs]))
cd = [mp.mpf(float(c[j])) for j in range(n + 1)]
worst = max(abs(sum(cd[j] * (b * i / 4000) ** j for j in range(n + 1)) - f(b * i / 4000)) for i in range(4001))
print("degree", n, "max |q - exact| with double coefficients:", mp.nstr(worst, 3), "-> relative error of log1p <=", mp.nstr(worst / 9, 3))
print(", ".join(float(c[j]).hex() for j in range(n, -1, -1)))
