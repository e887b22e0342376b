"""Coefficients of the odd polynomial atan(t) ~= t * Q(t^2) on |t| <= tan(pi/8) used by fdm_core.cuh (f16_atan_half).

Chebyshev interpolation of g(u) = atan(sqrt(u)) / sqrt(u) on u in [0, tan(pi/8)^2] in exact rational arithmetic (the series of g
converges like 0.1716^k there), converted to the monomial basis and rounded to double.  Prints the table and the maximum
error of the double-precision Horner evaluation against the rational reference."""
import math
from fractions import Fraction as F

DEG = 12                                   # degree of Q in u: 13 coefficients
T_MAX = F(41422, 100000)                   # a little above tan(pi/8) = 0.414213...
U_MAX = T_MAX * T_MAX


def g_exact(u: F, terms: int = 60) -> F:   # sum (-1)^k u^k / (2k+1)
    s, p = F(0), F(1)
    for k in range(terms):
        s += p / (2 * k + 1) if k % 2 == 0 else -p / (2 * k + 1)
        p *= u
    return s


def solve(A, b):                            # Gaussian elimination in Fractions
    n = len(b)
    M = [row[:] + [bi] for row, bi in zip(A, b)]
    for c in range(n):
        piv = max(range(c, n), key=lambda r: abs(M[r][c]))
        M[c], M[piv] = M[piv], M[c]
        for r in range(c + 1, n):
            f = M[r][c] / M[c][c]
            for k in range(c, n + 1):
                M[r][k] -= f * M[c][k]
    x = [F(0)] * n
    for r in range(n - 1, -1, -1):
        x[r] = (M[r][n] - sum(M[r][k] * x[k] for k in range(r + 1, n))) / M[r][r]
    return x


nodes = [U_MAX / 2 * (1 + F(math.cos(math.pi * (2 * i + 1) / (2 * (DEG + 1)))).limit_denominator(10 ** 18)) for i in range(DEG + 1)]
A = [[u ** k for k in range(DEG + 1)] for u in nodes]
coef = solve(A, [g_exact(u) for u in nodes])
cd = [float(c) for c in coef]
print("static constexpr double ATAN_Q[%d] = {" % (DEG + 1))
print(",\n".join("  %.17e" % c for c in cd) + "};")

worst = 0.0
N = 20001
for i in range(N):
    t = float(T_MAX) * (2 * i / (N - 1) - 1) * 0.99999
    u = t * t
    q = cd[DEG]
    for k in range(DEG - 1, -1, -1):
        q = q * u + cd[k]
    approx = t * q
    tf = F(t)
    exact = float(tf * g_exact(tf * tf))
    worst = max(worst, abs(approx - exact))
    assert abs(exact - math.atan(t)) < 3e-16
print("max |poly - atan| on the interval (double Horner, no fma): %.3e" % worst)
