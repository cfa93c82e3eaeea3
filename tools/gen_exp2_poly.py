#!/usr/bin/env python
"""Derives the polynomial used by csrc/sinkhorn_math.cuh: p(f) ~= 2^f on [-1/2, 1/2].
Chebyshev-node interpolation in 60-digit arithmetic (near-minimax), coefficients rounded to
double, error of the ROUNDED polynomial verified against mpmath on a dense grid.
    python tools/gen_exp2_poly.py [degree]
"""
import sys
import mpmath as mp

mp.mp.dps = 60


def fit(deg):
    n = deg + 1
    nodes = [mp.mpf(1) / 2 * mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x ** j
        b[i] = mp.mpf(2) ** x
    c = mp.lu_solve(A, b)
    coef = [float(c[j]) for j in range(n)]
    coef[0] = 1.0   # exact at f = 0
    return coef


def max_rel_err(coef, pts=4001):
    worst = mp.mpf(0)
    for i in range(pts):
        x = mp.mpf(-1) / 2 + mp.mpf(i) / (pts - 1)
        p = mp.mpf(0)
        for cj in reversed(coef):
            p = p * x + mp.mpf(cj)
        worst = max(worst, abs(p / mp.mpf(2) ** x - 1))
    return worst


if __name__ == "__main__":
    degs = [int(a) for a in sys.argv[1:]] or [10, 11, 12]
    for d in degs:
        c = fit(d)
        print(f"degree {d}: max rel err (exact arithmetic) = {mp.nstr(max_rel_err(c), 5)}")
        for j, cj in enumerate(c):
            print(f"    c{j} = {cj!r},  // {cj.hex()}")
