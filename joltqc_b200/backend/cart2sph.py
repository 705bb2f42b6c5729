"""Cartesian <-> real-spherical coefficient matrices, derived from the closed form.

The reference hard-codes libcint's ``normalized="sp"`` matrices in CUDA
(jqc/backend/common/cart2sph.cu:22-100).  Here they are generated from the explicit
solid-harmonic expansion (Helgaker, Jorgensen, Olsen, eq. 6.4.47) so that nothing is
transcribed:  c2s[l][cart, m] is the coefficient of the cartesian monomial (order: lx
descending, then ly descending — jqc/backend/util.py:21-36) in the real solid harmonic,
normalised as r^l Y_lm for l >= 2 and as plain x, y, z / 1 for l < 2 (the s/p angular
factors live in the contraction coefficients, jqc/pyscf/basis.py:547-553).
Spherical order: m = -l..l, except p which is (x, y, z) as in PySCF.
"""
from fractions import Fraction
from functools import lru_cache
from math import comb, factorial, pi, sqrt

import numpy as np


def cart_powers(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def _binom_half(n, k2):
    """binom(n, k) for integer k = k2/2 (k2 even)"""
    return comb(n, k2 // 2)


@lru_cache(maxsize=None)
def cart2sph_matrix(l):
    """(ncart, nsph) float64 matrix C with  phi_sph[m] = sum_c C[c, m] phi_cart[c]."""
    pw = cart_powers(l)
    idx = {p: n for n, p in enumerate(pw)}
    out = np.zeros((len(pw), 2 * l + 1))
    for m in range(-l, l + 1):
        am = abs(m)
        nlm = sqrt(2.0 * factorial(l + am) * factorial(l - am) / (2.0 if m == 0 else 1.0)) / (2**am * factorial(l))
        # v runs over integers for m >= 0 and half-integers for m < 0: use 2v
        v2_start = 0 if m >= 0 else 1
        for t in range((l - am) // 2 + 1):
            for u in range(t + 1):
                v2 = v2_start
                while v2 <= 2 * ((am - v2_start) // 2) + v2_start and v2 <= am:
                    sign = (-1) ** (t + (v2 - v2_start) // 2)
                    c = Fraction(sign, 4**t) * comb(l, t) * comb(l - t, am + t) * comb(t, u) * comb(am, v2)
                    ex = 2 * t + am - (2 * u + v2)
                    ey = 2 * u + v2
                    ez = l - 2 * t - am
                    out[idx[(ex, ey, ez)], m + l] += float(c) * nlm
                    v2 += 2
    if l >= 2:
        out *= sqrt((2 * l + 1) / (4.0 * pi))
    if l == 1:
        out = out[:, [2, 0, 1]]  # m = (-1, 0, 1) = (y, z, x)  ->  (x, y, z)
    return out


def nf_cart(l):
    return (l + 1) * (l + 2) // 2
