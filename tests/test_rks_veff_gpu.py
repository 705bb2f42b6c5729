"""RKS get_veff on the CUDA path: ``apply(ks)`` on an RKS-typed object installs get_jk / get_j / get_k
(CUDA engine) and the RKS glue (jqc/pyscf/rks.py:180-262, jqc/pyscf/__init__.py:186-214); the J/K parts
of the potential are compared with the CPU oracle for hybrid, pure and range-separated functionals, full
and incremental.  The XC quadrature is the host package's job (out of scope), so a linear stand-in
``nr_rks`` is used exactly as in tests/test_rks_veff_cpu.py.
"""
import numpy as np
import pytest

from tests.common import H2O, make, random_dm
from tests.test_rks_veff_cpu import _NumInt

pytestmark = pytest.mark.gpu


class _RKS:
    """The attributes of a GPU4PySCF RKS object that apply() and get_veff touch."""

    def __init__(self, mol, coeff):
        self.mol, self.xc, self.grids, self._eri, self.direct_scf = mol, "fake", None, None, True
        self.direct_scf_tol = 1e-13
        self._numint = _NumInt(*coeff)

    def istype(self, name):
        return name in ("RHF", "RKS")

    def do_nlc(self):
        return False

    # placeholders that apply() replaces
    def get_jk(self, *a, **k): raise AssertionError("apply() did not install get_jk")
    def get_j(self, *a, **k): raise AssertionError("apply() did not install get_j")
    def get_k(self, *a, **k): raise AssertionError("apply() did not install get_k")
    def get_veff(self, *a, **k): raise AssertionError("apply() did not install get_veff")

    def reset(self, mol=None):
        return self


def _np(x):
    x = getattr(x, "array", x)
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


@pytest.mark.parametrize("coeff", [(0.0, 0.0, 0.2), (0.0, 0.0, 0.0), (0.33, 0.65, 0.19)])
def test_rks_get_veff_through_apply(coeff):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from joltqc_b200.pyscf import apply
    from oracle.oracle import OracleJK
    mol, lay = make(H2O, "def2-tzvpp")
    ks = apply(_RKS(mol, coeff))
    assert ks._joltqc_applied and ks.get_veff.__func__.__module__ == "joltqc_b200.pyscf.rks"
    orc = OracleJK(lay)
    dm0, dm1 = random_dm(mol.nao, 1), random_dm(mol.nao, 2)
    omega, alpha, hyb = coeff

    def reference(dm):
        vj, vk = orc.get_jk(dm, 1, cutoff=1e-13)
        ref = 0.1 * dm + vj
        k = None
        if hyb or omega:
            k = hyb * vk
            if omega:
                k = k + (alpha - hyb) * orc.get_jk(dm, 1, with_j=False, omega=omega, cutoff=1e-13)[1]
            ref = ref - 0.5 * k
        return ref, vj, k

    v0 = ks.get_veff(mol, dm0)
    ref0, rj0, rk0 = reference(dm0)
    scale = max(1.0, np.abs(rj0).max())
    assert np.abs(_np(v0) - ref0).max() < 1e-10 * scale
    assert np.abs(_np(v0.vj) - rj0).max() < 1e-10 * scale
    if rk0 is not None:
        assert np.abs(_np(v0.vk) - rk0).max() < 1e-10 * scale
    else:
        assert v0.vk is None
    assert abs(float(v0.ecoul) - 0.5 * np.einsum("ij,ji", dm0, rj0)) < 1e-9 * scale
    # incremental build from (dm0, v0): J/K of the difference density only (jqc/pyscf/rks.py:226-250)
    v1 = ks.get_veff(mol, dm1, dm_last=dm0, vhf_last=v0)
    ref1, rj1, _ = reference(dm1)
    scale = max(1.0, np.abs(rj1).max())
    assert np.abs(_np(v1) - ref1).max() < 1e-10 * scale
    assert np.abs(_np(v1.vj) - rj1).max() < 1e-10 * scale
