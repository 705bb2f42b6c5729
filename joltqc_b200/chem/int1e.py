"""One-electron integrals (overlap, kinetic, nuclear attraction) on the host.

These are NOT on the J/K hot path; in a real deployment PySCF/GPU4PySCF supplies them.
They exist so that the stand-in RHF driver (chem/scf.py) can reproduce the reference's
golden total energies (jqc/pyscf/tests/test_scf.py:70,77) without PySCF.
McMurchie-Davidson scheme, cartesian components in the reference order (lx descending,
then ly descending), PySCF normalisation (s,p carry sqrt((2l+1)/4pi); l>=2 monomials are
not angularly normalised), spherical via backend.cart2sph.
"""
import math

import numpy as np
from scipy.special import hyp1f1

from ..backend.cart2sph import cart2sph_matrix, cart_powers
from . import mole as M


def _shells(mol):
    """Decontracted view: list of (atom_center, l, exps, coeffs[nctr, nprim]) per _bas row."""
    out = []
    for b in mol._bas:
        l, nprim, nctr = int(b[M.ANG_OF]), int(b[M.NPRIM_OF]), int(b[M.NCTR_OF])
        es = mol._env[b[M.PTR_EXP] : b[M.PTR_EXP] + nprim]
        cs = mol._env[b[M.PTR_COEFF] : b[M.PTR_COEFF] + nprim * nctr].reshape(nctr, nprim)
        fac = math.sqrt((2 * l + 1) / (4 * math.pi)) if l < 2 else 1.0
        r = mol._env[mol._atm[b[M.ATOM_OF], M.PTR_COORD] :][:3]
        out.append((r.copy(), l, es.copy(), cs * fac))
    return out


def _E(la, lb, a, b, xab):
    """Hermite expansion coefficients E[i, j, t] for one dimension."""
    p = a + b
    mu = a * b / p
    xpa = -b / p * xab
    xpb = a / p * xab
    E = np.zeros((la + 1, lb + 1, la + lb + 2))
    E[0, 0, 0] = math.exp(-mu * xab * xab)
    for i in range(la):
        for t in range(i + 2):
            v = xpa * E[i, 0, t] + (t + 1) * E[i, 0, t + 1]
            if t > 0:
                v += E[i, 0, t - 1] / (2 * p)
            E[i + 1, 0, t] = v
    for j in range(lb):
        for i in range(la + 1):
            for t in range(i + j + 2):
                v = xpb * E[i, j, t] + (t + 1) * E[i, j, t + 1]
                if t > 0:
                    v += E[i, j, t - 1] / (2 * p)
                E[i, j + 1, t] = v
    return E


def _R(tmax, p, pc):
    """Hermite Coulomb integrals R[t,u,v] (n = 0) up to t+u+v <= tmax."""
    T = p * float(pc @ pc)
    F = [hyp1f1(n + 0.5, n + 1.5, -T) / (2 * n + 1) for n in range(tmax + 1)]
    R = np.zeros((tmax + 1, tmax + 1, tmax + 1, tmax + 1))  # [n,t,u,v]
    for n in range(tmax + 1):
        R[n, 0, 0, 0] = (-2 * p) ** n * F[n]
    for s in range(1, tmax + 1):
        for n in range(tmax + 1 - s):
            for t in range(s + 1):
                for u in range(s - t + 1):
                    v = s - t - u
                    if t > 0:
                        val = pc[0] * R[n + 1, t - 1, u, v]
                        if t > 1:
                            val += (t - 1) * R[n + 1, t - 2, u, v]
                    elif u > 0:
                        val = pc[1] * R[n + 1, t, u - 1, v]
                        if u > 1:
                            val += (u - 1) * R[n + 1, t, u - 2, v]
                    else:
                        val = pc[2] * R[n + 1, t, u, v - 1]
                        if v > 1:
                            val += (v - 1) * R[n + 1, t, u, v - 2]
                    R[n, t, u, v] = val
    return R[0]


def _pair_block(sa, sb, charges, centers):
    ra, la, ea, ca = sa
    rb, lb, eb, cb = sb
    pa, pb = cart_powers(la), cart_powers(lb)
    S = np.zeros((ca.shape[0], cb.shape[0], len(pa), len(pb)))
    T = np.zeros_like(S)
    V = np.zeros_like(S)
    ab = ra - rb
    for ia, a in enumerate(ea):
        for ib, b in enumerate(eb):
            p = a + b
            P = (a * ra + b * rb) / p
            E = [_E(la, lb + 2, a, b, ab[d]) for d in range(3)]
            pref = (math.pi / p) ** 1.5
            Rs = [(z, _R(la + lb, p, P - c)) for z, c in zip(charges, centers)]
            s = np.zeros((len(pa), len(pb)))
            t = np.zeros_like(s)
            v = np.zeros_like(s)
            for m, (ax, ay, az) in enumerate(pa):
                for n, (bx, by, bz) in enumerate(pb):
                    sx, sy, sz = E[0][ax, bx, 0], E[1][ay, by, 0], E[2][az, bz, 0]
                    s[m, n] = sx * sy * sz * pref

                    def kin(Ed, i, j):
                        val = -2 * b * (2 * j + 1) * Ed[i, j, 0] + 4 * b * b * Ed[i, j + 2, 0]
                        if j >= 2:
                            val += j * (j - 1) * Ed[i, j - 2, 0]
                        return -0.5 * val

                    t[m, n] = (kin(E[0], ax, bx) * sy * sz + sx * kin(E[1], ay, by) * sz
                               + sx * sy * kin(E[2], az, bz)) * pref
                    acc = 0.0
                    ex = E[0][ax, bx, : ax + bx + 1]
                    ey = E[1][ay, by, : ay + by + 1]
                    ez = E[2][az, bz, : az + bz + 1]
                    w = np.einsum("t,u,v->tuv", ex, ey, ez)
                    for z, R in Rs:
                        acc -= z * float((w * R[: ax + bx + 1, : ay + by + 1, : az + bz + 1]).sum())
                    v[m, n] = acc * 2 * math.pi / p
            cc = np.einsum("i,j->ij", ca[:, ia], cb[:, ib])
            S += cc[:, :, None, None] * s
            T += cc[:, :, None, None] * t
            V += cc[:, :, None, None] * v
    return S, T, V


def int1e(mol):
    """Returns (S, T, V) in the molecule's AO basis (cartesian or real-spherical)."""
    sh = _shells(mol)
    charges = mol.atom_charges()
    centers = mol.atom_coords()
    loc = mol.ao_loc_nr(cart=True)
    n = int(loc[-1])
    S = np.zeros((n, n))
    T = np.zeros((n, n))
    V = np.zeros((n, n))
    for i, sa in enumerate(sh):
        for j in range(i + 1):
            sb = sh[j]
            s, t, v = _pair_block(sa, sb, charges, centers)
            nca, ncb = s.shape[0], s.shape[1]
            na, nb = s.shape[2], s.shape[3]
            for m, out in ((s, S), (t, T), (v, V)):
                blk = m.transpose(0, 2, 1, 3).reshape(nca * na, ncb * nb)
                out[loc[i] : loc[i + 1], loc[j] : loc[j + 1]] = blk
                out[loc[j] : loc[j + 1], loc[i] : loc[i + 1]] = blk.T
    if mol.cart:
        return S, T, V
    C = cart2sph_total(mol)
    return C.T @ S @ C, C.T @ T @ C, C.T @ V @ C


def cart2sph_total(mol):
    """(nao_cart, nao_sph) block-diagonal transformation for the molecule's own shells."""
    lc, ls = mol.ao_loc_nr(cart=True), mol.ao_loc_nr(cart=False)
    C = np.zeros((int(lc[-1]), int(ls[-1])))
    for ib, b in enumerate(mol._bas):
        l, nctr = int(b[M.ANG_OF]), int(b[M.NCTR_OF])
        c = cart2sph_matrix(l)
        for k in range(nctr):
            C[lc[ib] + k * c.shape[0] : lc[ib] + (k + 1) * c.shape[0],
              ls[ib] + k * c.shape[1] : ls[ib] + (k + 1) * c.shape[1]] = c
    return C
