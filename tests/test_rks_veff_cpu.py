"""RKS get_veff glue (jqc/pyscf/rks.py:180-262) on CPU with the oracle as get_jk and a fake
numerical-integration object: hybrid, pure and range-separated functionals, full vs incremental."""
import numpy as np
import pytest

from joltqc_b200.pyscf.rks import generate_get_veff
from tests.common import H2O, make, random_dm


class _LibXC:
    def __init__(self, hybrid): self.hybrid = hybrid
    def is_hybrid_xc(self, xc): return self.hybrid
    def is_nlc(self, xc): return False


class _NumInt:
    def __init__(self, omega, alpha, hyb):
        self.coeff = (omega, alpha, hyb)
        self.libxc = _LibXC(hyb != 0 or omega != 0)
    def nr_rks(self, mol, grids, xc, dm):
        return 10.0, -1.5, 0.1 * np.asarray(dm)          # any linear "vxc" will do
    def rsh_and_hybrid_coeff(self, xc, spin=0): return self.coeff


class _KS:
    def __init__(self, mol, orc, coeff):
        self.mol, self.xc, self.grids, self._eri, self.direct_scf = mol, "fake", None, None, True
        self._numint = _NumInt(*coeff)
        self.get_jk = lambda m, dm, hermi=1, omega=None: orc.get_jk(dm, hermi, omega=omega)
        self.get_j = lambda m, dm, hermi=1: orc.get_jk(dm, hermi, with_k=False)[0]
        self.get_k = lambda m, dm, hermi=1, omega=None: orc.get_jk(dm, hermi, with_j=False, omega=omega)[1]
    def do_nlc(self): return False


@pytest.mark.parametrize("coeff", [(0.0, 0.0, 0.2), (0.0, 0.0, 0.0), (0.33, 0.65, 0.19)])
def test_rks_get_veff(coeff):
    from oracle.oracle import OracleJK
    mol, lay = make(H2O, "def2-svp")
    orc = OracleJK(lay)
    ks = _KS(mol, orc, coeff)
    get_veff = generate_get_veff()
    dm0, dm1 = random_dm(mol.nao, 1), random_dm(mol.nao, 2)
    omega, alpha, hyb = coeff
    v0 = get_veff(ks, mol, dm0)
    vj, vk = orc.get_jk(dm0, 1)
    ref = 0.1 * dm0 + vj
    if hyb or omega:
        k = hyb * vk
        if omega:
            k = k + (alpha - hyb) * orc.get_jk(dm0, 1, with_j=False, omega=omega)[1]
        ref = ref - 0.5 * k
        assert np.abs(np.asarray(v0.vk) - k).max() < 1e-10
    assert np.abs(np.asarray(v0) - ref).max() < 1e-10
    assert abs(v0.ecoul - 0.5 * np.einsum("ij,ji", dm0, vj)) < 1e-9
    # incremental build from (dm0, v0) must equal the full build at dm1
    v1_inc = get_veff(ks, mol, dm1, dm_last=dm0, vhf_last=v0)
    v1_full = get_veff(ks, mol, dm1)
    assert np.abs(np.asarray(v1_inc) - np.asarray(v1_full)).max() < 1e-9
    assert np.abs(np.asarray(v1_inc.vj) - np.asarray(v1_full.vj)).max() < 1e-9
