"""Shared helpers for the parity tests (CPU oracle vs CUDA engine)."""
import numpy as np

from joltqc_b200.chem.mole import M
from joltqc_b200.pyscf.basis import BasisLayout

H2O = """
O       0.0000000000    -0.0000000000     0.1174000000
H      -0.7570000000    -0.0000000000    -0.4696000000
H       0.7570000000     0.0000000000    -0.4696000000
"""

# same molecule as jqc/pyscf/tests/test_jk.py:57-66 (H2 in Bohr)
H2_BOHR = """
H  -0.757    4.   -0.4696
H   0.757    4.   -0.4696
"""


def benzene():
    rc, rh = 1.39, 1.39 + 1.09
    out = []
    for k in range(6):
        a = np.pi / 3 * k
        out.append(("C", (rc * np.cos(a), rc * np.sin(a), 0.0)))
    for k in range(6):
        a = np.pi / 3 * k
        out.append(("H", (rh * np.cos(a), rh * np.sin(a), 0.0)))
    return out


def make(atom, basis, cart=False, unit="Angstrom"):
    mol = M(atom=atom, basis=basis, cart=cart, unit=unit)
    return mol, BasisLayout.from_mol(mol, alignment=4)


def random_dm(nao, seed=9, n=None, symmetric=True):
    """Seeded dm = R R^T as in jqc/pyscf/tests/test_jk.py:68-71."""
    rng = np.random.RandomState(seed)
    shape = (nao, nao) if n is None else (n, nao, nao)
    dm = rng.rand(*shape)
    if symmetric:
        dm = dm @ dm.swapaxes(-1, -2) if n is None else np.einsum("bij,bkj->bik", dm, dm)
    return dm
