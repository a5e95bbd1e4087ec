#!/usr/bin/env python
"""Generate Gauss-Kronrod node/weight tables (QUADPACK layout) with mpmath.

Test infrastructure only.  QUADPACK / GSL `gsl_integration_qag` (key 1..6 = 15/21/31/41/51/61
points) is a third-party dependency of the reference (cosmology.c:389, hmf.c:628) that is not
vendored under /root/reference, so the oracle restates the published algorithm.  The node tables
are derived here from first principles instead of being typed in:

  * Gauss nodes: roots of Legendre P_n.
  * Kronrod nodes: roots of the Stieltjes polynomial E_{n+1}, defined by
        int_{-1}^{1} P_n(x) E_{n+1}(x) x^k dx = 0   for k = 0..n,
    expanded in the Legendre basis (Patterson 1968 / Piessens & Branders 1974).
  * Weights: exactness on P_k, k = 0..3n+1 (solved as a linear system at 60 digits).

Output layout follows QUADPACK: xgk[0..n] descending positive abscissae (xgk[n] = 0), wgk the
Kronrod weights, wg the Gauss weights for the embedded n-point rule (only the positive half).
"""
import sys
from mpmath import mp, mpf, legendre, matrix, lu_solve, findroot, quad, polyroots

mp.dps = 60


def gauss_nodes(n):
    # roots of P_n via Newton from Chebyshev-like guesses
    import numpy as np
    xs = []
    for x0 in np.polynomial.legendre.leggauss(n)[0]:
        x = mpf(float(x0))
        for _ in range(6):  # Newton polish; P_n'(x) = n (x P_n - P_{n-1}) / (x^2 - 1)
            pn = legendre(n, x)
            dp = n * (x * pn - legendre(n - 1, x)) / (x * x - 1)
            x = x - pn / dp
        xs.append(x)
    xs = sorted(xs)
    if n % 2:
        xs[n // 2] = mpf(0)
    return xs


def bracket_root(f, lo, hi):
    """Bisection to ~1e-20 then Newton polish (central-difference derivative at 60 digits)."""
    flo, fhi = f(lo), f(hi)
    assert flo * fhi < 0, "Stieltjes root not bracketed"
    for _ in range(70):
        mid = (lo + hi) / 2
        fm = f(mid)
        if flo * fm <= 0:
            hi, fhi = mid, fm
        else:
            lo, flo = mid, fm
    x = (lo + hi) / 2
    h = mpf(10) ** (-25)
    for _ in range(4):
        d = (f(x + h) - f(x - h)) / (2 * h)
        x = x - f(x) / d
    return x


def triple(n, j, k, gx, gw):
    return sum(w * legendre(n, x) * legendre(j, x) * legendre(k, x) for x, w in zip(gx, gw))


def gauss_rule(m):
    xs = gauss_nodes(m)
    ws = []
    for x in xs:
        dp = m * (x * legendre(m, x) - legendre(m - 1, x)) / (x * x - 1)
        ws.append(2 / ((1 - x * x) * dp * dp))
    return xs, ws


def kronrod(n):
    """Return (xk, wk, xg, wg) for the (2n+1)-point Kronrod extension of n-point Gauss."""
    xg, wg = gauss_rule(n)
    # high order rule to integrate triple products exactly: degree <= n + (n+1) + n
    hx, hw = gauss_rule((3 * n + 3) // 2 + 2)
    # E_{n+1} = P_{n+1} + sum_{j in J} c_j P_j, J = {n-1, n-3, ...} (same parity as n+1)
    J = list(range(n - 1, -1, -2))
    # test functions P_k: P_n * E_{n+1} is odd, so only odd k <= n give non-trivial conditions
    K = list(range(1, n + 1, 2))
    assert len(K) == len(J)
    A = matrix(len(K), len(J))
    b = matrix(len(K), 1)
    for r, k in enumerate(K):
        for c, j in enumerate(J):
            A[r, c] = triple(n, j, k, hx, hw)
        b[r] = -triple(n, n + 1, k, hx, hw)
    cj = lu_solve(A, b)

    def E(x):
        return legendre(n + 1, x) + sum(cj[i] * legendre(j, x) for i, j in enumerate(J))

    # Kronrod-only nodes interlace with the Gauss nodes (and +-1)
    bounds = [mpf(-1)] + xg + [mpf(1)]
    xe = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        xe.append(bracket_root(E, lo, hi))
    xk = sorted(xg + xe)
    m = len(xk)
    # weights from exactness on P_0..P_{m-1}
    V = matrix(m, m)
    rhs = matrix(m, 1)
    for k in range(m):
        for i, x in enumerate(xk):
            V[k, i] = legendre(k, x)
        rhs[k] = 2 if k == 0 else 0
    wk = lu_solve(V, rhs)
    return xk, [wk[i] for i in range(m)], xg, wg


def fmt(v):
    return mp.nstr(v, 36, strip_zeros=False)


def emit(n, out):
    xk, wk, xg, wg = kronrod(n)
    m = 2 * n + 1
    # positive half, descending: index 0 = largest abscissa, index n = 0
    xs = [xk[m - 1 - i] for i in range(n + 1)]
    ws = [wk[m - 1 - i] for i in range(n + 1)]
    xs[n] = mpf(0)
    # gauss weights for positive gauss nodes, descending (QUADPACK: wg[j] pairs with xgk[2j+1])
    ng = (n + 1) // 2
    wgs = [wg[n - 1 - i] for i in range(ng)]
    # sanity: rule integrates x^(2n) .. exactly up to degree 3n+1
    for deg in (0, 2, 2 * n, 3 * n + 1 - ((3 * n + 1) % 2)):
        s = sum(w * x ** deg for x, w in zip(xk, wk))
        assert abs(s - mpf(2) / (deg + 1)) < mpf(10) ** (-45), (n, deg, s)
    out.write(f"static const double GK{m}_XGK[{n + 1}] = {{\n")
    out.write(",\n".join("    " + fmt(x) for x in xs) + "};\n")
    out.write(f"static const double GK{m}_WGK[{n + 1}] = {{\n")
    out.write(",\n".join("    " + fmt(w) for w in ws) + "};\n")
    out.write(f"static const double GK{m}_WG[{ng}] = {{\n")
    out.write(",\n".join("    " + fmt(w) for w in wgs) + "};\n\n")


if __name__ == "__main__":
    path = sys.argv[1]
    with open(path, "w") as f:
        f.write("/* Generated by tools/gen_gk_tables.py (mpmath, 60 digits). Do not edit. */\n")
        f.write("/* Gauss-Kronrod abscissae/weights in QUADPACK layout (positive half, descending). */\n\n")
        for n in (7, 10, 15, 20, 25, 30):
            emit(n, f)
    print("wrote", path)
