"""A minimal stand-in for ``pyscf.gto.Mole`` (PySCF is absent from the build/GPU images).

Only what the J/K path consumes is provided: the libcint-layout ``_atm/_bas/_env`` integer
and float tables (column meanings cited from the reference's use of them,
jqc/pyscf/basis.py:519-528, 757-766), ``ao_loc``, ``nao``, ``cart``, coordinates, charges
and the nuclear repulsion energy.  Contraction coefficients are normalised exactly like
PySCF's ``make_bas_env`` (radial ``gto_norm`` followed by normalisation of the contracted
function), because the reference kernels consume ``_env`` coefficients as-is
(jqc/pyscf/basis.py:581-584).
"""
from __future__ import annotations

import math
import re
import sys

import numpy as np

from . import basis_data

# libcint slots
ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, KAPPA_OF, PTR_EXP, PTR_COEFF, RESERVE_BASLOT = range(8)
BAS_SLOTS = 8
CHARGE_OF, PTR_COORD, NUC_MOD_OF, PTR_ZETA, PTR_FRAC_CHARGE, RESERVE_ATMSLOT = range(6)
ATM_SLOTS = 6
PTR_ENV_START = 20
BOHR = 0.52917721092  # PySCF's nist.BOHR (CODATA 2010/2014 rounding used by pyscf.data.nist)

ELEMENTS = ["X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar"]
ANGULAR = "spdfghik"


def gaussian_int(n, alpha):
    """int_0^inf r^n exp(-alpha r^2) dr"""
    n1 = (n + 1) * 0.5
    return math.gamma(n1) / (2.0 * alpha**n1)


def gto_norm(l, expnt):
    """Radial normalisation of r^l exp(-a r^2) (same definition as pyscf.gto.gto_norm)."""
    return 1.0 / math.sqrt(gaussian_int(l * 2 + 2, 2.0 * expnt))


def _normalize_contracted(l, es, cs):
    ee = es[:, None] + es[None, :]
    ee = np.vectorize(lambda a: gaussian_int(l * 2 + 2, a))(ee)
    s1 = 1.0 / np.sqrt(np.einsum("pi,pq,qi->i", cs, ee, cs))
    return cs * s1[None, :]


def parse_atoms(atom, unit="Angstrom"):
    """Accepts a PySCF-style string ('O 0 0 0; H 0 1 0' or newline separated) or a list of
    (symbol, (x, y, z)).  Returns symbols and coordinates in Bohr."""
    out = []
    if isinstance(atom, str):
        for line in re.split(r"[;\n]", atom):
            tok = line.replace(",", " ").split()
            if not tok:
                continue
            out.append((tok[0], tuple(float(v) for v in tok[1:4])))
    else:
        for sym, xyz in atom:
            out.append((sym, tuple(float(v) for v in xyz)))
    symbols = [re.sub(r"[^A-Za-z]", "", s).capitalize() for s, _ in out]
    coords = np.array([c for _, c in out], dtype=np.float64).reshape(-1, 3)
    if not unit.upper().startswith(("B", "AU")):
        coords = coords / BOHR
    return symbols, coords


def read_xyz(path):
    with open(path) as f:
        lines = f.read().splitlines()
    n = int(lines[0].split()[0])
    return "\n".join(lines[2 : 2 + n])


class Mole:
    def __init__(self, atom=None, basis="sto-3g", unit="Angstrom", cart=False, charge=0, spin=0,
                 verbose=0, output=None, max_memory=4000, **kw):
        self.atom = atom
        self.basis = basis
        self.unit = unit
        self.cart = bool(cart)
        self.charge = charge
        self.spin = spin
        self.verbose = verbose
        self.output = output
        self.max_memory = max_memory
        self.stdout = sys.stdout
        self._atm = np.zeros((0, ATM_SLOTS), dtype=np.int32)
        self._bas = np.zeros((0, BAS_SLOTS), dtype=np.int32)
        self._env = np.zeros(PTR_ENV_START)
        self._built = False

    # ------------------------------------------------------------------ build
    def build(self):
        symbols, coords = parse_atoms(self.atom, self.unit)
        self._symbols = symbols
        env = [0.0] * PTR_ENV_START
        atm = []
        for sym, xyz in zip(symbols, coords):
            ptr = len(env)
            env.extend(xyz.tolist())
            env.append(0.0)  # zeta slot
            atm.append([ELEMENTS.index(sym), ptr, 1, ptr + 3, 0, 0])
        bas = []
        cache = {}
        for ia, sym in enumerate(symbols):
            if sym not in cache:
                shells = []
                for l, prims in sorted(self._basis_for(sym), key=lambda s: s[0]):
                    arr = np.asarray(prims, dtype=np.float64)
                    es = arr[:, 0].copy()
                    cs = arr[:, 1:].copy()
                    cs = cs * np.array([gto_norm(l, e) for e in es])[:, None]
                    cs = _normalize_contracted(l, es, cs)
                    ptr_exp = len(env)
                    env.extend(es.tolist())
                    ptr_coeff = len(env)
                    env.extend(cs.T.reshape(-1).tolist())  # contraction-major
                    shells.append((l, es.size, cs.shape[1], ptr_exp, ptr_coeff))
                cache[sym] = shells
            for l, nprim, nctr, ptr_exp, ptr_coeff in cache[sym]:
                bas.append([ia, l, nprim, nctr, 0, ptr_exp, ptr_coeff, 0])
        self._atm = np.asarray(atm, dtype=np.int32).reshape(-1, ATM_SLOTS)
        self._bas = np.asarray(bas, dtype=np.int32).reshape(-1, BAS_SLOTS)
        self._env = np.asarray(env, dtype=np.float64)
        self._built = True
        return self

    def _basis_for(self, sym):
        if isinstance(self.basis, dict):
            b = self.basis.get(sym, self.basis.get("default"))
            if isinstance(b, str):
                return basis_data.load(b, sym)
            return b
        return basis_data.load(self.basis, sym)

    # ------------------------------------------------------------- properties
    @property
    def natm(self):
        return int(self._atm.shape[0])

    @property
    def nbas(self):
        return int(self._bas.shape[0])

    def atom_coords(self):
        ptr = self._atm[:, PTR_COORD]
        return np.stack([self._env[p : p + 3] for p in ptr]) if len(ptr) else np.zeros((0, 3))

    def atom_charges(self):
        return self._atm[:, CHARGE_OF].astype(np.float64)

    def atom_symbol(self, i):
        return ELEMENTS[int(self._atm[i, CHARGE_OF])]

    @property
    def nelectron(self):
        return int(self._atm[:, CHARGE_OF].sum()) - self.charge

    def _dims(self, cart=None):
        cart = self.cart if cart is None else cart
        l = self._bas[:, ANG_OF]
        per = (l + 1) * (l + 2) // 2 if cart else 2 * l + 1
        return per * self._bas[:, NCTR_OF]

    @property
    def ao_loc(self):
        return self.ao_loc_nr()

    def ao_loc_nr(self, cart=None):
        loc = np.zeros(self.nbas + 1, dtype=np.int32)
        np.cumsum(self._dims(cart), out=loc[1:])
        return loc

    @property
    def nao(self):
        return int(self._dims().sum())

    def nao_nr(self):
        return self.nao

    def energy_nuc(self):
        z = self.atom_charges()
        r = self.atom_coords()
        e = 0.0
        for i in range(self.natm):
            for j in range(i):
                e += z[i] * z[j] / np.linalg.norm(r[i] - r[j])
        return e

    def copy(self):
        m = Mole(self.atom, self.basis, self.unit, self.cart, self.charge, self.spin, self.verbose)
        m._atm, m._bas, m._env = self._atm.copy(), self._bas.copy(), self._env.copy()
        m._symbols = list(getattr(self, "_symbols", []))
        m._built = self._built
        return m

    def set_geom_(self, atom, unit="Angstrom"):
        self.atom, self.unit = atom, unit
        return self.build()


def M(**kw):
    """Same convenience constructor as ``pyscf.M``."""
    return Mole(**kw).build()
