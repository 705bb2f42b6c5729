"""GPU parity against the UNMODIFIED reference kernels (oracle/_ref cubins compiled from
/root/reference by oracle/ref_kernels/build_ref_kernels.py and launched through cuda-python):
element-wise comparison of the kernel-side J/K accumulators of the engine with what
rys_1q1t_vjk / rys_1qnt_vjk / screen_jk_tasks of JoltQC produce on the same shell table, density
and cutoffs (north_star: max-abs 1e-10).  Skipped when the cubins were not built."""
import numpy as np
import pytest

from tests.common import H2O, benzene, make, random_dm

pytestmark = pytest.mark.gpu


def _compare(atom, basis, seed, scale=1.0, tol=1e-10, mol=None):
    import torch
    from oracle.ref_kernels import runner
    if not runner.available():
        pytest.skip("oracle/_ref not built (needs the reference tree at build time)")
    if mol is None:
        mol, lay = make(atom, basis)
    else:
        from joltqc_b200.pyscf.basis import BasisLayout
        lay = BasisLayout.from_mol(mol, alignment=4)
    eng = lay.engine()
    ref = runner.RefJK(lay, eng)
    missing = ref.missing_kernels()
    if missing:
        pytest.skip("oracle/_ref lacks %d kernels for this basis (e.g. %s)" % (len(missing), missing[0]))
    dm = random_dm(mol.nao, seed) * scale
    dk = eng.dm_from_mol(dm)
    rj, rk = ref.get_jk_raw(dk, hermi=1, cutoff=1e-13)
    buf = eng.build_partial(dm, hermi=1)
    torch.cuda.synchronize()
    n2 = lay.nao * lay.nao
    vj, vk = buf[:n2].reshape(lay.nao, lay.nao), buf[n2:2 * n2].reshape(lay.nao, lay.nao)
    ej = (vj - rj).abs().max().item() / max(1.0, rj.abs().max().item())
    ek = (vk - rk).abs().max().item() / max(1.0, rk.abs().max().item())
    counts, _, _ = eng.last_stats()
    assert int(counts.sum()) == ref.last["quartets"], (int(counts.sum()), ref.last["quartets"])
    assert ej < tol and ek < tol, (ej, ek)


def test_h2o_tzvpp_vs_reference_kernels():
    _compare(H2O, "def2-tzvpp", 9)


def test_benzene_ccpvtz_vs_reference_kernels():
    _compare(benzene(), "cc-pvtz", 9, scale=1.0 / 264)


def test_taxol_svp_vs_reference_kernels():
    """BASELINE config 3 (taxol stand-in / def2-SVP, 1e9 quartets): a molecule large enough that every
    brick / multi-lane launch runs many tasks per warp, element-wise against the reference's kernels"""
    import bench
    mol, _ = bench.build_mol("taxol-svp")
    _compare(None, None, 11, scale=1.0 / mol.nao, mol=mol)
