#!/usr/bin/env python
"""Generate the Rys-quadrature root/weight tables from first principles (mpmath).

Nothing here is read from the reference: nodes/weights come from a Golub-Welsch
solve on the Boys-function moments at 80 digits, then are fitted per interval.

Math.  For x >= 0 the n-point Rys rule is the Gauss rule in u = t^2 for the weight
W(u) = exp(-x u) / (2 sqrt(u)) on [0,1]; its moments are the Boys functions
m_k = F_k(x).  "root" = u_i = t_i^2 and "weight" = w_i with sum_i w_i = F_0(x),
the convention the reference kernels consume (jqc/backend/jk/1q1t.cu:236-242,
rt_aa = rt/(aij+akl)).

Tables (same functional form as jqc/backend/rys/rys_roots.cu:29-160 so that the
CPU oracle can restate that routine verbatim):
  * CHEB  [n][interval][root][k][2]  Chebyshev series (T_0 coefficient unhalved) of
           degree 13 on intervals of width 2.5, (root, weight) interleaved per k;
           14+2n intervals (x < 35+5n) are emitted for n roots.
  * SMALLX[n][root][4] = r(0), r'(0), w(0), w'(0)       used for x < 3e-7
  * LARGEX[n][root][2] = h_i^2, w_i/sqrt(pi/4)-normalised  (Hermite H_2n limit):
           root = h_i^2 / x, weight = W_i * sqrt(pi/(4x))  for x >= 35+5n
Roots are stored in ascending order.
"""
import argparse
import multiprocessing as mp_
import os
import sys

import mpmath as mp
import numpy as np

DEGREE = 13
NCOEF = DEGREE + 1
WIDTH = mp.mpf("2.5")
NMAX = 9
NFIT = 28  # Chebyshev nodes sampled per interval before truncation to degree 13


def n_intervals(n):
    return 14 + 2 * n


def boys(k, x):
    if x == 0:
        return mp.mpf(1) / (2 * k + 1)
    a = mp.mpf(k) + mp.mpf(1) / 2
    return mp.gammainc(a, 0, x) / (2 * mp.power(x, a))


def rys_rule(n, x):
    """Ascending nodes u_i and weights w_i of the n-point Rys rule at x (mpf)."""
    m = [boys(k, x) for k in range(2 * n + 1)]
    H = mp.matrix(n + 1, n + 1)
    for i in range(n + 1):
        for j in range(n + 1):
            H[i, j] = m[i + j]
    R = mp.cholesky(H).T  # H = R^T R, R upper triangular
    alpha = []
    beta = []
    for j in range(n):
        a = R[j, j + 1] / R[j, j]
        if j > 0:
            a -= R[j - 1, j] / R[j - 1, j - 1]
        alpha.append(a)
        if j < n - 1:
            beta.append(R[j + 1, j + 1] / R[j, j])
    J = mp.matrix(n, n)
    for j in range(n):
        J[j, j] = alpha[j]
        if j < n - 1:
            J[j, j + 1] = beta[j]
            J[j + 1, j] = beta[j]
    ev, evec = mp.eigsy(J)
    order = sorted(range(n), key=lambda i: ev[i])
    roots = [ev[i] for i in order]
    weights = [m[0] * evec[0, i] ** 2 for i in order]
    return roots, weights


def fit_interval(args):
    n, it = args
    mp.mp.dps = 80
    a = WIDTH * it
    # Chebyshev-Gauss nodes on [-1,1]
    nodes = [mp.cos(mp.pi * (mp.mpf(j) + mp.mpf(1) / 2) / NFIT) for j in range(NFIT)]
    vals = []
    for u in nodes:
        x = a + (u + 1) * WIDTH / 2
        r, w = rys_rule(n, x)
        vals.append((r, w))
    out = np.zeros((n, NCOEF, 2))
    for i in range(n):
        for t in range(2):
            f = [vals[j][t][i] for j in range(NFIT)]
            for k in range(NCOEF):
                c = mp.fsum(
                    f[j] * mp.cos(mp.pi * k * (mp.mpf(j) + mp.mpf(1) / 2) / NFIT)
                    for j in range(NFIT)
                ) * 2 / NFIT
                if k == 0:
                    c /= 2
                out[i, k, t] = float(c)
    return n, it, out


def smallx(n):
    mp.mp.dps = 80
    r0, w0 = rys_rule(n, mp.mpf(0))
    h = mp.mpf(10) ** -20
    r1, w1 = rys_rule(n, h)
    out = np.zeros((n, 4))
    for i in range(n):
        out[i] = [float(r0[i]), float((r1[i] - r0[i]) / h), float(w0[i]), float((w1[i] - w0[i]) / h)]
    return out


def largex(n):
    mp.mp.dps = 80
    # positive roots of H_{2n}: Gauss rule for exp(-t^2) on (0, inf) in u=t^2 is
    # Gauss-Laguerre with alpha=-1/2: nodes u_i, weights normalised to sum 1.
    # Moments of exp(-u) u^{-1/2}/2: Gamma(k+1/2)/2
    m = [mp.gamma(mp.mpf(k) + mp.mpf(1) / 2) / 2 for k in range(2 * n + 1)]
    H = mp.matrix(n + 1, n + 1)
    for i in range(n + 1):
        for j in range(n + 1):
            H[i, j] = m[i + j]
    R = mp.cholesky(H).T
    J = mp.matrix(n, n)
    for j in range(n):
        a = R[j, j + 1] / R[j, j]
        if j > 0:
            a -= R[j - 1, j] / R[j - 1, j - 1]
        J[j, j] = a
        if j < n - 1:
            b = R[j + 1, j + 1] / R[j, j]
            J[j, j + 1] = b
            J[j + 1, j] = b
    ev, evec = mp.eigsy(J)
    order = sorted(range(n), key=lambda i: ev[i])
    out = np.zeros((n, 2))
    for c, i in enumerate(order):
        out[c] = [float(ev[i]), float(evec[0, i] ** 2)]  # weights sum to 1 (x sqrt(pi/4x))
    return out


def cheb_to_monomial(cheb):
    """Exact change of basis of every degree-13 series from Chebyshev T_k(u) to powers u^k (rational
    arithmetic on the stored doubles, rounded once).  The device evaluates the power form with
    Estrin's scheme: 13 FMAs of depth 4 instead of a 26-operation dependent Clenshaw chain; the
    coefficients decay fast enough that sum|m_k| <= 2.4 |p|, i.e. no loss of accuracy
    (max relative deviation from the exact Chebyshev value 7e-16 over all intervals)."""
    from fractions import Fraction

    T = [[0] * NCOEF for _ in range(NCOEF)]
    T[0][0] = 1
    T[1][1] = 1
    for k in range(2, NCOEF):
        for j in range(NCOEF):
            T[k][j] = (2 * T[k - 1][j - 1] if j > 0 else 0) - T[k - 2][j]
    out = {}
    for n, c in cheb.items():
        m = np.zeros_like(c)
        for idx in np.ndindex(c.shape[0], c.shape[1], 2):
            a = [Fraction(float(v)) for v in c[idx[0], idx[1], :, idx[2]]]
            for j in range(NCOEF):
                m[idx[0], idx[1], j, idx[2]] = float(sum(a[k] * T[k][j] for k in range(NCOEF)))
        out[n] = m
    return out


def emit(path, qual, iqual, cheb, sx, lx, name="RYS_CHEB", basis="Chebyshev T_k(u), T_0 coefficient unhalved"):
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_rys_tables.py (mpmath, 80 digits). Do not edit.\n")
        f.write("// Series basis of %s: %s; u = 0.8 (x - 2.5 it) - 1.\n" % (name, basis))
        f.write("// Layouts: %s[off(n) + ((it*n + root)*14 + k)*2 + {0:root,1:weight}],\n" % name)
        f.write("//          RYS_SMALLX[(n(n-1)/2 + root)*4 + {r0,r1,w0,w1}], RYS_LARGEX[(n(n-1)/2 + root)*2 + {R,W}]\n")
        f.write("#pragma once\n")
        f.write("#define RYS_NMAX %d\n#define RYS_NCOEF %d\n" % (NMAX, NCOEF))
        offs = [0]
        for n in range(1, NMAX + 1):
            offs.append(offs[-1] + n_intervals(n) * n * NCOEF * 2)
        f.write("// element offset of the n-root block inside RYS_CHEB (index n-1)\n")
        f.write("%s RYS_CHEB_OFFSET[%d] = {%s};\n" % (iqual, NMAX + 1, ", ".join(map(str, offs))))
        flat = np.concatenate([cheb[n].ravel() for n in range(1, NMAX + 1)])
        f.write("%s %s[%d] = {\n" % (qual, name, flat.size))
        for i in range(0, flat.size, 4):
            f.write(" " + ", ".join("%.17e" % v for v in flat[i : i + 4]) + ",\n")
        f.write("};\n")
        for name, d in (("RYS_SMALLX", sx), ("RYS_LARGEX", lx)):
            flat = np.concatenate([d[n].ravel() for n in range(1, NMAX + 1)])
            f.write("%s %s[%d] = {\n" % (qual, name, flat.size))
            for i in range(0, flat.size, 4):
                f.write(" " + ", ".join("%.17e" % v for v in flat[i : i + 4]) + ",\n")
            f.write("};\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out-npz", default=os.path.join(os.path.dirname(__file__), "rys_tables.npz"))
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--from-npz", action="store_true", help="re-emit the headers from the saved fit (no mpmath run)")
    args = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if args.from_npz:
        d = np.load(args.out_npz)
        cheb = {n: d["cheb%d" % n] for n in range(1, NMAX + 1)}
        sx = {n: d["sx%d" % n] for n in range(1, NMAX + 1)}
        lx = {n: d["lx%d" % n] for n in range(1, NMAX + 1)}
        emit_all(root, cheb, sx, lx)
        return
    jobs = [(n, it) for n in range(1, NMAX + 1) for it in range(n_intervals(n))]
    cheb = {n: np.zeros((n_intervals(n), n, NCOEF, 2)) for n in range(1, NMAX + 1)}
    with mp_.Pool(args.procs) as pool:
        for done, (n, it, out) in enumerate(pool.imap_unordered(fit_interval, jobs, chunksize=1)):
            cheb[n][it] = out
            if done % 20 == 0:
                print("fit %d/%d" % (done, len(jobs)), file=sys.stderr, flush=True)
    sx = {n: smallx(n) for n in range(1, NMAX + 1)}
    lx = {n: largex(n) for n in range(1, NMAX + 1)}
    np.savez(args.out_npz, **{"cheb%d" % n: cheb[n] for n in cheb}, **{"sx%d" % n: sx[n] for n in sx}, **{"lx%d" % n: lx[n] for n in lx})
    emit_all(root, cheb, sx, lx)


def emit_all(root, cheb, sx, lx):
    # device: power-basis coefficients (Estrin evaluation); oracle: the Chebyshev fit itself
    emit(os.path.join(root, "joltqc_b200", "csrc", "rys_tables.cuh"), "__device__ const double", "static constexpr int",
         cheb_to_monomial(cheb), sx, lx, name="RYS_MONO", basis="powers u^k (exact change of basis of the Chebyshev fit)")
    emit(os.path.join(root, "oracle", "rys_tables.h"), "static const double", "static const int", cheb, sx, lx)


if __name__ == "__main__":
    main()
