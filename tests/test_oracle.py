"""CPU tests pinning the oracle (no GPU): reference golden vectors and independent math."""
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle
from tests.common import H2O, make, random_dm

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_rys_roots_match_reference_tables():
    """Roots/weights vs values evaluated from the reference's own tables
    (tests/golden/make_rys_golden.py <- jqc/backend/rys/rys_root{1..9}.cu)."""
    g = json.load(open(os.path.join(GOLD, "rys_ref_samples.json")))
    worst = 0.0
    for n, rows in g.items():
        for x, r, w in rows:
            rr, ww = oracle.rys_roots(int(n), x)
            o = np.argsort(rr)
            worst = max(worst, np.abs(rr[o] - r).max(), np.abs(ww[o] - w).max())
    assert worst < 1e-12, worst


def test_rys_rule_integrates_boys_moments():
    """Independent check: an n-point rule must reproduce F_k(x), k < 2n (Boys moments)."""
    from scipy.special import hyp1f1
    for n in (1, 2, 3, 5, 7, 9):
        for x in (0.0, 1e-7, 0.3, 2.5, 11.7, 33.3, 60.0, 200.0):
            r, w = oracle.rys_roots(n, x)
            for k in range(2 * n):
                fk = hyp1f1(k + 0.5, k + 1.5, -x) / (2 * k + 1)
                assert abs(np.sum(w * r**k) - fk) < 2e-13 * max(1.0, fk), (n, x, k)


def test_ssss_closed_form():
    """(ss|ss) of four normalised s primitives against the textbook Boys-function formula."""
    from joltqc_b200.chem.mole import M
    from joltqc_b200.pyscf.basis import BasisLayout
    mol = M(atom="H 0 0 0; H 0 0 1.4; H 0.3 1.1 0.2; H -1 0.4 0.9", unit="B",
            basis={"H": [(0, [(0.9, 1.0)]), (0, [(0.35, 1.0)])]})
    lay = BasisLayout.from_mol(mol)
    orc = oracle.OracleJK(lay)
    from scipy.special import erf
    rec = lay.basis_data_fp64["packed"]
    rng = np.random.RandomState(1)
    idx = np.nonzero(~lay.pad_id)[0]
    for _ in range(20):
        i, j, k, l = rng.choice(idx, 4)
        val = orc.eri_block(i, j, k, l)[0, 0, 0, 0]
        a, b, c, d = (rec[s, 5] for s in (i, j, k, l))
        A, B, C, D = (rec[s, :3] for s in (i, j, k, l))
        p, q = a + b, c + d
        P, Q = (a * A + b * B) / p, (c * C + d * D) / q
        T = p * q / (p + q) * np.sum((P - Q) ** 2)
        f0 = 1.0 if T < 1e-14 else 0.5 * math.sqrt(math.pi / T) * erf(math.sqrt(T))
        norm = np.prod([(2 * e / math.pi) ** 0.75 for e in (a, b, c, d)])
        ref = (norm * 2 * math.pi**2.5 / (p * q * math.sqrt(p + q)) * math.exp(-a * b / p * np.sum((A - B) ** 2))
               * math.exp(-c * d / q * np.sum((C - D) ** 2)) * f0)
        assert abs(val - ref) < 1e-13 * max(1, abs(ref))


def test_eri_permutational_symmetry():
    mol, lay = make(H2O, "def2-tzvpp")
    orc = oracle.OracleJK(lay)
    idx = np.nonzero(~lay.pad_id)[0]
    rng = np.random.RandomState(3)
    for _ in range(12):
        i, j, k, l = rng.choice(idx, 4)
        a = orc.eri_block(i, j, k, l)
        assert np.abs(a - orc.eri_block(j, i, l, k).transpose(1, 0, 3, 2)).max() < 1e-12
        assert np.abs(a - orc.eri_block(k, l, i, j).transpose(2, 3, 0, 1)).max() < 1e-12


def test_jk_against_dense_eri_contraction():
    """get_jk of the oracle (screened, 8-fold symmetric, split shells) equals the brute-force
    contraction of the full ERI tensor built block by block — general contraction (cc-pVTZ C),
    nprim > 3 splitting and the AO transforms included."""
    for basis, cart in (("cc-pvtz", False), ("def2-svp", True)):
        mol, lay = make("C 0 0 0; H 0 0 1.1", basis, cart=cart)
        orc = oracle.OracleJK(lay)
        T = orc.transform()
        nao = lay.nao
        loc = lay.ao_loc
        eri = np.zeros((nao,) * 4)
        idx = np.nonzero(~lay.pad_id)[0]
        for i in idx:
            for j in idx:
                for k in idx:
                    for l in idx:
                        eri[loc[i]:loc[i + 1], loc[j]:loc[j + 1], loc[k]:loc[k + 1], loc[l]:loc[l + 1]] = \
                            orc.eri_block(i, j, k, l)
        dm = random_dm(mol.nao, seed=5)
        dmi = T @ dm @ T.T
        j_ref = T.T @ np.einsum("ijkl,lk->ij", eri, dmi) @ T
        k_ref = T.T @ np.einsum("ijkl,jk->il", eri, dmi) @ T
        vj, vk = orc.get_jk(dm, hermi=1)
        assert np.abs(vj - j_ref).max() < 1e-10
        assert np.abs(vk - k_ref).max() < 1e-10
        dm0 = random_dm(mol.nao, seed=6, symmetric=False)
        dmi = T @ dm0 @ T.T
        j_ref = T.T @ np.einsum("ijkl,lk->ij", eri, dmi) @ T
        k_ref = T.T @ np.einsum("ijkl,jk->il", eri, dmi) @ T
        vj, vk = orc.get_jk(dm0, hermi=0)
        assert np.abs(vj - j_ref).max() < 1e-10
        assert np.abs(vk - k_ref).max() < 1e-10


@pytest.mark.parametrize("cart,e_ref", [(False, -76.0624634523), (True, -76.0627443874)])
def test_reference_golden_rhf_energy(cart, e_ref):
    """Known-answer test of the whole restatement: RHF H2O/def2-TZVPP total energies quoted in
    jqc/pyscf/tests/test_scf.py:70,77 (reference tolerance 1e-5; the oracle reaches 1e-9)."""
    from joltqc_b200.chem.scf import RHF
    mol, lay = make(H2O, "def2-tzvpp", cart=cart)
    orc = oracle.OracleJK(lay)
    mf = RHF(mol)
    mf.conv_tol = 1e-10
    mf.get_jk = lambda m=None, dm=None, hermi=1, **kw: orc.get_jk(dm, hermi, cutoff=mf.direct_scf_tol)
    e = mf.kernel()
    assert mf.converged
    assert abs(e - e_ref) < 1e-9, e - e_ref


def test_oracle_reproduces_golden_jk_fixture():
    """tests/golden/jk_small_cases.npz (frozen by tests/golden/make_jk_golden.py): the oracle of today reproduces the
    committed J/K vectors — a regression pin for the checker itself."""
    import os
    from oracle.oracle import OracleJK
    from tests.golden.make_jk_golden import CASES
    from tests.common import make
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jk_small_cases.npz"))
    for name, (atom, basis, cart, unit, seed, hermi, omega) in CASES.items():
        mol, lay = make(atom, basis, cart=cart, unit=unit)
        vj, vk = OracleJK(lay).get_jk(g[name + "_dm"], hermi, True, True, omega, 1e-13)
        assert np.abs(vj - g[name + "_vj"]).max() < 1e-11 * np.abs(vj).max(), name
        assert np.abs(vk - g[name + "_vk"]).max() < 1e-11 * np.abs(vk).max(), name
