"""GPU parity tests: CUDA engine (through the C ABI) vs the CPU oracle on the same inputs.

Mirrors jqc/pyscf/tests/test_jk.py:57-276, test_basis_sets_jk.py:29-91 and test_scf.py:67-108.
Bar (north_star): max-abs elementwise 1e-10 on J and K (the reference's own tests only ask 1e-7
against libcint), total SCF energy 1e-9 Ha.
"""
import numpy as np
import pytest

from tests.common import H2_BOHR, H2O, benzene, make, random_dm

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch


def _oracle(lay):
    from oracle.oracle import OracleJK
    return OracleJK(lay)


def _check(lay, dm, hermi, with_j=True, with_k=True, omega=None, tol=TOL, cutoff=1e-13):
    from joltqc_b200.pyscf.jk import generate_jk_kernel
    get_jk = generate_jk_kernel(lay, cutoff_fp64=cutoff, cutoff_fp32=cutoff)
    vj, vk = get_jk(lay._mol, dm, hermi=hermi, with_j=with_j, with_k=with_k, omega=omega)
    rj, rk = _oracle(lay).get_jk(dm, hermi, with_j, with_k, omega, cutoff)
    if with_j:
        vj = vj.cpu().numpy()
        assert vj.shape == np.asarray(dm).shape
        scale = max(1.0, np.abs(rj).max())
        assert np.abs(vj - rj).max() < tol * scale, ("J", np.abs(vj - rj).max())
    else:
        assert isinstance(vj, int) and vj == 0
    if with_k:
        vk = vk.cpu().numpy()
        scale = max(1.0, np.abs(rk).max())
        assert np.abs(vk - rk).max() < tol * scale, ("K", np.abs(vk - rk).max())
    else:
        assert isinstance(vk, int) and vk == 0
    return get_jk


@pytest.mark.parametrize("cart", [False, True])
def test_h2_tzvpp_fp64(torch_cuda, cart):
    mol, lay = make(H2_BOHR, "def2-tzvpp", cart=cart, unit="B")
    _check(lay, random_dm(mol.nao, 9), hermi=1)


def test_multiple_dms(torch_cuda):
    mol, lay = make(H2_BOHR, "def2-tzvpp", unit="B")
    _check(lay, random_dm(mol.nao, 9, n=3), hermi=1)


def test_j_only_k_only(torch_cuda):
    mol, lay = make(H2O, "def2-tzvpp")
    dm = random_dm(mol.nao, 9)
    _check(lay, dm, hermi=1, with_k=False)
    _check(lay, dm, hermi=1, with_j=False)


def test_hermi0_nonsymmetric(torch_cuda):
    mol, lay = make(H2O, "def2-tzvpp")
    _check(lay, random_dm(mol.nao, 11, symmetric=False), hermi=0)
    _check(lay, random_dm(mol.nao, 12, n=2, symmetric=False), hermi=0)


def test_omega_long_range(torch_cuda):
    mol, lay = make(H2O, "def2-tzvpp")
    _check(lay, random_dm(mol.nao, 9), hermi=1, omega=0.5)
    with pytest.raises(AssertionError):
        from joltqc_b200.pyscf.jk import generate_jk_kernel
        generate_jk_kernel(lay)(mol, random_dm(mol.nao, 9), hermi=1, omega=-0.3)


def test_far_apart_atoms_screening(torch_cuda):
    """two atoms 100 Bohr apart: whole classes are screened out (test_jk.py:250-276)"""
    mol, lay = make("H 0 0 0; H 0 0 100.0", "def2-tzvpp", unit="B")
    _check(lay, random_dm(mol.nao, 9), hermi=1)


@pytest.mark.parametrize("basis", ["sto-3g", "def2-svp", "def2-tzvp", "def2-tzvpp"])
def test_h2o_basis_sets(torch_cuda, basis):
    mol, lay = make(H2O, basis)
    _check(lay, random_dm(mol.nao, 42), hermi=1)


def test_general_contraction_ccpvtz(torch_cuda):
    mol, lay = make("C 0 0 0; H 0 0 1.09; H 1.03 0 -0.36; H -0.51 0.89 -0.36; H -0.51 -0.89 -0.36", "cc-pvtz")
    _check(lay, random_dm(mol.nao, 42), hermi=1)


def test_g_functions(torch_cuda):
    mol, lay = make(H2O, "test-spdfg")
    _check(lay, random_dm(mol.nao, 7), hermi=1)
    mol, lay = make(H2O, "test-spdfg", cart=True)
    _check(lay, random_dm(mol.nao, 7), hermi=1)


def test_loose_cutoff_same_quartet_set(torch_cuda):
    """with a loose threshold many quartets are dropped; the engine must drop the same ones"""
    mol, lay = make(benzene(), "def2-svp")
    dm = random_dm(mol.nao, 3) * 1e-3
    _check(lay, dm, hermi=1, cutoff=1e-7, tol=1e-12)
    eng = lay.engine()
    counts, _, _ = eng.last_stats()
    orc = _oracle(lay)
    orc.get_jk(dm, 1, True, True, None, 1e-7)
    assert int(counts.sum()) == int(orc.last_nquartets)
    assert np.array_equal(counts, orc.last_counts)


def test_benzene_ccpvtz(torch_cuda):
    """BASELINE.json config 2"""
    mol, lay = make(benzene(), "cc-pvtz")
    assert mol.nao == 264
    _check(lay, random_dm(mol.nao, 9) / mol.nao, hermi=1)


def test_q_matrix_and_transforms(torch_cuda):
    for cart in (False, True):
        mol, lay = make(H2O, "test-spdfg", cart=cart)
        eng = lay.engine()
        orc = _oracle(lay)
        for omega in (0.0, 0.4):
            q = eng.q_matrix(omega).cpu().numpy()
            qr = orc.q_matrix(omega)
            assert np.abs(q - qr).max() < 2e-5, np.abs(q - qr).max()   # float32 logs
        T = orc.transform()
        dm = random_dm(mol.nao, 4, n=2, symmetric=False)
        di = eng.dm_from_mol(dm).cpu().numpy()
        assert np.abs(di - np.stack([T @ d @ T.T for d in dm])).max() < 1e-12
        vi = np.random.RandomState(2).rand(2, lay.nao, lay.nao)
        vm = eng.dm_to_mol(vi).cpu().numpy()
        assert np.abs(vm - np.stack([T.T @ v @ T for v in vi])).max() < 1e-11


def test_host_entry_point(torch_cuda):
    mol, lay = make(H2O, "def2-tzvp")
    dm = random_dm(mol.nao, 5)
    vj, vk = lay.engine().get_jk_host(dm, hermi=1)
    rj, rk = _oracle(lay).get_jk(dm, 1)
    assert np.abs(vj - rj).max() < TOL * np.abs(rj).max() and np.abs(vk - rk).max() < TOL * np.abs(rk).max()


def test_run_to_run_noise_floor(torch_cuda):
    """FP64 atomics make the summation order vary: two runs must agree far below the 1e-10 bar."""
    mol, lay = make(benzene(), "def2-svp")
    dm = random_dm(mol.nao, 1)
    eng = lay.engine()
    a = eng.get_jk(dm, hermi=1)
    b = eng.get_jk(dm, hermi=1)
    for x, y in zip(a, b):
        assert (x - y).abs().max().item() < 1e-11 * x.abs().max().item()


@pytest.mark.parametrize("cart,e_ref", [(False, -76.0624634523), (True, -76.0627443874)])
def test_scf_golden_energy_through_apply(torch_cuda, cart, e_ref):
    """jqc/pyscf/tests/test_scf.py:67-79 through the plugin entry point"""
    import joltqc_b200.pyscf as jq
    from joltqc_b200.chem.scf import RHF
    mol, _ = make(H2O, "def2-tzvpp", cart=cart)
    mf = RHF(mol)
    mf.conv_tol = 1e-10
    mf = jq.apply(mf)
    e = mf.kernel()
    assert mf.converged and abs(e - e_ref) < 1e-9, e - e_ref


def test_apply_reset_and_errors(torch_cuda):
    import joltqc_b200.pyscf as jq
    from joltqc_b200.chem.scf import RHF
    mol, _ = make(H2O, "sto-3g")
    mf = jq.apply(RHF(mol))
    assert mf._joltqc_applied
    e1 = mf.kernel()
    mol2, _ = make("O 0 0 0.13; H -0.76 0 -0.47; H 0.76 0 -0.47", "sto-3g")
    mf2 = mf.reset(mol2)
    assert mf2._joltqc_applied and mf2._jqc_layout._mol is mol2
    e2 = mf2.kernel()
    assert abs(e1 - e2) > 1e-6
    with pytest.raises(ValueError):
        mf2.get_jk(mol2, np.zeros((3, 3)))


def test_shards_sum_to_full_build(torch_cuda):
    """jqc_engine_set_shard / jqc_build_partial / jqc_finalize on ONE GPU: the partial [J||K]
    buffers of rank 0..2 of a world of 3, summed on the device (what the NCCL all_reduce does),
    finalise to the same J and K as the unsharded build."""
    import ctypes
    import torch
    from joltqc_b200.backend import lib as _lib
    from joltqc_b200.backend.engine import JKEngine, _wrap_device_buffer
    mol, lay = make(benzene(), "def2-svp")
    dm = random_dm(mol.nao, 5)
    full_j, full_k = lay.engine().get_jk(dm, hermi=1)
    eng = JKEngine(lay)
    d3 = torch.as_tensor(dm, device=eng.device).reshape(1, mol.nao, mol.nao).contiguous()
    total = None
    counts = 0
    for rank in range(3):
        eng.set_shard(rank, 3)
        p, ln = ctypes.c_void_p(), ctypes.c_size_t()
        _lib.check(eng.L.jqc_build_partial(eng.h, d3.data_ptr(), 1, 1, 1, 1, 0.0, 1e-13, 1e-13, ctypes.byref(p),
                                           ctypes.byref(ln), None))
        buf = _wrap_device_buffer(p.value, ln.value, eng.device)
        torch.cuda.synchronize()
        counts += int(eng.last_stats()[0].sum())
        total = buf.clone() if total is None else total + buf
    buf.copy_(total)                       # stand-in for all_reduce(SUM) into the engine's buffer
    vj, vk = torch.empty_like(d3), torch.empty_like(d3)
    _lib.check(eng.L.jqc_finalize(eng.h, vj.data_ptr(), vk.data_ptr(), None))
    torch.cuda.synchronize()
    assert (vj[0] - full_j).abs().max().item() < 1e-11 * full_j.abs().max().item()
    assert (vk[0] - full_k).abs().max().item() < 1e-11 * full_k.abs().max().item()
    assert counts == int(lay.engine().last_stats()[0].sum())      # every quartet exactly once


@pytest.mark.parametrize("seed", range(6))
def test_randomized_parity(torch_cuda, seed):
    """random small molecules / bases / densities / options against the oracle"""
    rng = np.random.RandomState(100 + seed)
    elems = ["H", "C", "N", "O"]
    basis = ["sto-3g", "def2-svp", "def2-tzvp", "def2-tzvpp"][rng.randint(4)]
    natm = rng.randint(2, 5)
    atom = [(elems[rng.randint(4)], tuple(rng.uniform(-2.2, 2.2, 3))) for _ in range(natm)]
    # keep atoms apart
    xyz = np.array([a[1] for a in atom])
    for a in range(natm):
        for b in range(a):
            if np.linalg.norm(xyz[a] - xyz[b]) < 0.7:
                xyz[a] += 1.5
    atom = [(atom[a][0], tuple(xyz[a])) for a in range(natm)]
    cart = bool(rng.randint(2))
    mol, lay = make(atom, basis, cart=cart)
    hermi = int(rng.randint(2))
    n = None if rng.randint(2) else 2
    dm = random_dm(mol.nao, 200 + seed, n=n, symmetric=bool(hermi))
    omega = None if rng.randint(3) else 0.3
    with_j, with_k = [(True, True), (True, False), (False, True)][rng.randint(3)]
    cutoff = [1e-13, 1e-10][rng.randint(2)]
    _check(lay, dm, hermi=hermi, with_j=with_j, with_k=with_k, omega=omega, cutoff=cutoff)


def test_python_partial_finalize_api(torch_cuda):
    """JKEngine.build_partial / finalize (the two-phase multi-GPU API) on a single rank"""
    mol, lay = make(H2O, "def2-svp")
    dm = random_dm(mol.nao, 8)
    eng = lay.engine()
    ref_j, ref_k = eng.get_jk(dm, hermi=1)
    buf = eng.build_partial(dm, hermi=1)
    assert buf.numel() == 2 * lay.nao * lay.nao
    vj, vk = eng.finalize()
    assert (vj - ref_j).abs().max().item() < 1e-11 and (vk - ref_k).abs().max().item() < 1e-11


# ---------------------------------------------------------------------------------------
# Engine paths selected by environment knobs (read at engine creation): every parity case
# below builds a fresh layout -> fresh engine.

def _fresh_check(monkeypatch, env, atom, basis, seed=9, scale=1.0, **kw):
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))
    mol, lay = make(atom, basis)
    _check(lay, random_dm(mol.nao, seed) * scale, hermi=1, **kw)
    return lay


def test_quartet_list_path_small_classes(torch_cuda, monkeypatch):
    """JQC_BRICK=0: the one-quartet-per-thread kernels fed by screen_tasks_kernel (round-1 path,
    still used for n_dm > 1) against the oracle."""
    _fresh_check(monkeypatch, {"JQC_BRICK": 0}, H2O, "def2-tzvpp")
    _fresh_check(monkeypatch, {"JQC_BRICK": 0}, benzene(), "def2-svp", scale=0.05)


def test_tile16_path(torch_cuda, monkeypatch):
    """JQC_SMALL_TILES=1: the 4x4-tile variant of the small-class kernel (jk_tile16.cuh)."""
    _fresh_check(monkeypatch, {"JQC_SMALL_TILES": 1}, H2O, "def2-tzvpp")
    _fresh_check(monkeypatch, {"JQC_SMALL_TILES": 1}, benzene(), "def2-svp", scale=0.05)


@pytest.mark.parametrize("brick", [0, 1])
def test_multi_chunk_task_queues(torch_cuda, monkeypatch, brick):
    """Shrunk queue capacity / kl chunk: benzene/cc-pVTZ then runs with many (ij, kl) chunks per
    group quartet, i.e. the ij0 > 0 / kl0 > 0 iterations of the engine's chunk loops that only a
    valinomycin-sized molecule reaches with the production sizes."""
    monkeypatch.setenv("JQC_BRICK", str(brick))
    mol, lay0 = make(benzene(), "cc-pvtz")
    dm = random_dm(mol.nao, 9) / 264
    lay0.engine().get_jk(dm, hermi=1)
    _, _, launches0 = lay0.engine().last_stats()
    lay = _fresh_check(monkeypatch, {"JQC_QUEUE_CAP": 4096, "JQC_KL_CHUNK": 5}, benzene(), "cc-pvtz", scale=1.0 / 264)
    _, _, launches = lay.engine().last_stats()
    if brick == 0:     # (with the brick kernels on, no class of this molecule uses the queue any more)
        assert launches > 3 * launches0, (launches, launches0)     # many more chunks were really launched


@pytest.mark.parametrize("ichunk", [1, 3, 64])
def test_brick_chunking(torch_cuda, monkeypatch, ichunk):
    """brick kernel with different bra-chunk sizes (task decomposition must not change the result)"""
    lay = _fresh_check(monkeypatch, {"JQC_BRICK_ICHUNK": ichunk}, benzene(), "def2-tzvp", scale=0.02)
    counts, _, _ = lay.engine().last_stats()
    orc = _oracle(lay)
    orc.get_jk(random_dm(lay._mol.nao, 9) * 0.02, 1)
    assert np.array_equal(counts, orc.last_counts)


def test_brick_density_screening_counts(torch_cuda):
    """brick path, density-dominated screening: same quartets per class as the oracle, J-only and
    K-only variants use their own density criterion (screen_jk_tasks.cu:241-261)"""
    mol, lay = make(benzene(), "def2-tzvp")
    rng = np.random.RandomState(5)
    dm = rng.randn(mol.nao, mol.nao) * np.exp(-rng.uniform(0, 14, (mol.nao, mol.nao)))
    dm = dm + dm.T
    for with_j, with_k in [(True, True), (True, False), (False, True)]:
        _check(lay, dm, hermi=1, with_j=with_j, with_k=with_k, cutoff=1e-9, tol=1e-12)
        counts, _, _ = lay.engine().last_stats()
        orc = _oracle(lay)
        orc.get_jk(dm, 1, with_j, with_k, None, 1e-9)
        assert np.array_equal(counts, orc.last_counts), (with_j, with_k)


@pytest.mark.parametrize("mode", [0, 2])
def test_multilane_kernel_variants(torch_cuda, monkeypatch, mode):
    """JQC_BWARP=0: every large class on the quartet-list multi-lane kernel (jk_warp.cuh);
    JQC_BWARP=2: every large class on the brick-scheduled one (jk_bwarp.cuh).  The default (1)
    mixes them by a measured table, so both have to agree with the oracle on their own."""
    lay = _fresh_check(monkeypatch, {"JQC_BWARP": mode}, benzene(), "cc-pvtz", scale=1.0 / 264)
    counts, _, _ = lay.engine().last_stats()
    orc = _oracle(lay)
    orc.get_jk(random_dm(lay._mol.nao, 9) / 264, 1)
    assert np.array_equal(counts, orc.last_counts)
    _fresh_check(monkeypatch, {"JQC_BWARP": mode}, H2O, "def2-tzvpp", omega=0.3)


def test_engine_matches_golden_fixture(torch_cuda):
    """The CUDA engine against the committed golden J/K vectors (tests/golden/jk_small_cases.npz) — the same
    comparison as the oracle tests above, but with nothing of oracle/ executed."""
    import os
    from joltqc_b200.pyscf.jk import generate_jk_kernel
    from tests.golden.make_jk_golden import CASES
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jk_small_cases.npz"))
    for name, (atom, basis, cart, unit, seed, hermi, omega) in CASES.items():
        mol, lay = make(atom, basis, cart=cart, unit=unit)
        vj, vk = generate_jk_kernel(lay)(mol, g[name + "_dm"], hermi=hermi, omega=omega)
        rj, rk = g[name + "_vj"], g[name + "_vk"]
        assert np.abs(vj.cpu().numpy() - rj).max() < TOL * max(1.0, np.abs(rj).max()), name
        assert np.abs(vk.cpu().numpy() - rk).max() < TOL * max(1.0, np.abs(rk).max()), name
