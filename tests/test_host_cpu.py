"""CPU-only tests: host logic, the C-ABI library surface, the FLOP model (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.common import H2O, make

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """libjoltqc_b200.so must load and export exactly what include/joltqc_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "joltqc_b200.h")).read()
    declared = set(re.findall(r"\b(jqc_[a-z0-9_]+)\s*\(", hdr))
    from joltqc_b200.backend import lib
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    L = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly (no silent fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import joltqc_b200.pyscf as jq
    from joltqc_b200.chem.scf import RHF
    mol, lay = make(H2O, "sto-3g")
    with pytest.raises(RuntimeError):
        jq.apply(RHF(mol))
    with pytest.raises(NotImplementedError):
        RHF(mol).get_jk(mol, np.eye(mol.nao))
    # the C entry point itself refuses as well
    from joltqc_b200.backend import lib
    L = lib.load()
    h = ctypes.c_void_p()
    packed = np.ascontiguousarray(lay.basis_data_fp64["packed"])
    angs = np.ascontiguousarray(lay.angs, dtype=np.int32)
    nprims = np.ascontiguousarray(lay.nprims, dtype=np.int32)
    ao_loc = np.ascontiguousarray(lay.ao_loc, dtype=np.int32)
    pad = np.ascontiguousarray(lay.pad_id, dtype=np.uint8)
    goff = np.ascontiguousarray(lay.group_offset, dtype=np.int32)
    moff = np.ascontiguousarray(lay.mol_ao_offset, dtype=np.int32)
    p = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    d = lib.BasisDesc(nbas=angs.size, records=p(packed, ctypes.c_double), angs=p(angs, ctypes.c_int),
                      nprims=p(nprims, ctypes.c_int), ao_loc=p(ao_loc, ctypes.c_int), pad=p(pad, ctypes.c_uint8),
                      ngroups=goff.size - 1, group_offset=p(goff, ctypes.c_int), mol_ao_offset=p(moff, ctypes.c_int),
                      mol_nao=lay.mol_nao, mol_cart=0, c2s=None)
    rc = L.jqc_engine_create(ctypes.byref(d), 0, ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in L.jqc_last_error()


def test_cart2sph_formula_matches_libcint_tables():
    """closed-form matrices (product) vs the numeric libcint tables kept by the oracle
    (reference: jqc/backend/common/cart2sph.cu:22-100; jqc/backend/tests/test_cart2sph.py)."""
    from joltqc_b200.backend.cart2sph import cart2sph_matrix
    from oracle.oracle import c2s_matrix
    for l in range(5):
        assert np.abs(cart2sph_matrix(l) - c2s_matrix(l)).max() < 1e-14


def test_cart2sph_orthonormal_overlap():
    """cart2sph(S_cart) of one shell is the identity for normalised spherical GTOs."""
    from joltqc_b200.chem import int1e
    from joltqc_b200.chem.mole import M
    mol = M(atom="O 0 0 0", basis="test-spdfg")
    s, _, _ = int1e.int1e(mol)
    loc = mol.ao_loc
    for ib in range(mol.nbas):
        blk = s[loc[ib]:loc[ib + 1], loc[ib]:loc[ib + 1]]
        assert np.abs(blk - np.eye(blk.shape[0])).max() < 1e-10


@pytest.mark.parametrize("basis,nshell,nao_cart", [("def2-tzvpp", 25, 70), ("def2-svp", 13, 26), ("sto-3g", 5, 7)])
def test_layout_sizes(basis, nshell, nao_cart):
    """H2O/def2-TZVPP -> 25 split shells, 70 kernel-side cartesian AOs."""
    mol, lay = make(H2O, basis)
    assert int((~lay.pad_id).sum()) == nshell
    assert lay.nao == nao_cart


def test_layout_invariants():
    """Mirrors jqc/pyscf/tests/test_basis_layout.py: grouping, padding, nprim limit, maps."""
    from joltqc_b200.constants import NPRIM_MAX, TILE
    for basis, atom in (("cc-pvtz", "C 0 0 0; H 0 0 1.1"), ("def2-tzvp", H2O)):
        mol, lay = make(atom, basis)
        assert lay.nbasis % TILE == 0 and np.all(lay.group_offset % TILE == 0)
        assert lay.nprims.max() <= NPRIM_MAX
        keys = [tuple(k) for k in lay.group_key]
        assert keys == sorted(keys, key=lambda k: (k[0], -k[1])) and len(set(keys)) == len(keys)
        for g, (l, npr) in enumerate(keys):
            s0, s1 = lay.group_offset[g], lay.group_offset[g + 1]
            assert np.all(lay.angs[s0:s1] == l) and np.all(lay.nprims[s0:s1] == npr)
            pads = lay.pad_id[s0:s1]
            assert not pads[0] and np.all(np.diff(pads.astype(int)) >= 0)      # pads trail the group
            assert np.all(lay.ce[s0:s1][pads] == lay.ce[s0]) if pads.any() else True
        ao = lay.ao_loc
        widths = np.diff(ao)
        assert np.all(widths[lay.pad_id] == 0)
        assert np.all(widths[~lay.pad_id] == ((lay.angs + 1) * (lay.angs + 2) // 2)[~lay.pad_id])
        # every molecular AO is covered, split siblings alias the same AOs
        moff = lay.mol_ao_offset
        assert np.all(moff[lay.pad_id] == -1) and moff[~lay.pad_id].min() == 0
        assert lay.mol_nao == mol.nao
        # s/p coefficients carry sqrt((2l+1)/4pi) relative to the _env values (basis.py:547-553)
        sm = lay.splitted_mol
        for s in np.nonzero(~lay.pad_id)[0][:8]:
            b = sm._bas[lay.to_split_map[s]]
            fac = np.sqrt((2 * b[1] + 1) / (4 * np.pi)) if b[1] < 2 else 1.0
            assert np.allclose(lay.ce[s, 0], sm._env[b[6]] * fac)
            assert np.allclose(lay.ce[s, 1], sm._env[b[5]])


def test_general_contraction_is_decontracted():
    """cc-pVTZ carbon: one 8-primitive s block with 2 contractions -> 2 x (3+3+2) split shells."""
    mol, lay = make("C 0 0 0", "cc-pvtz")
    s_shells = int(((lay.angs == 0) & ~lay.pad_id).sum())
    assert s_shells == 2 * 3 + 2
    assert mol.nao == 30


def test_flop_model_matches_survey():
    """SURVEY 8(d) check values: (pp|pp) P=1 -> 2511 flop; (dd|dd) J+K -> 24570 + 15552."""
    import bench
    key = lambda a, b, c, d: ((a * 5 + b) * 5 + c) * 5 + d
    fe, fd = bench.flops_per_class(key(1, 1, 1, 1))
    assert fe + fd == 2511
    fe, fd = bench.flops_per_class(key(2, 2, 2, 2))
    assert (fe, fd) == (24570, 15552)


def test_apply_keeps_reference_surface():
    """apply()/get_default_config() expose the reference's names and config keys
    (jqc/pyscf/__init__.py:100-121)."""
    import joltqc_b200.pyscf as jq
    cfg = jq.get_default_config()
    assert set(cfg["jk"]) == {"cutoff_fp32", "cutoff_fp64"} and "dft" in cfg
    import inspect
    from joltqc_b200.pyscf import jk
    sig = inspect.signature(jk.generate_jk_kernel)
    assert list(sig.parameters) == ["basis_layout", "cutoff_fp64", "cutoff_fp32"]
    for name in ("generate_get_j", "generate_get_k", "generate_get_jk", "generate_get_veff"):
        assert hasattr(jk, name)


def test_jk_pair_surface(monkeypatch):
    """joltqc_b200.pyscf.jk_pair mirrors jqc/pyscf/jk_pair.py:49-115 (names, argument order incl. pair_wide_vk)
    and hands the cutoffs to the same engine-backed generator."""
    import inspect
    from joltqc_b200.pyscf import jk, jk_pair
    seen = []
    monkeypatch.setattr(jk, "generate_jk_kernel",
                        lambda lay, cutoff_fp64=1e-13, cutoff_fp32=1e-13: seen.append((lay, cutoff_fp64, cutoff_fp32)) or
                        (lambda *a, **k: (("J", k), ("K", k))))
    for name in ("generate_get_j", "generate_get_k", "generate_get_jk", "generate_jk_kernel"):
        sig = inspect.signature(getattr(jk_pair, name))
        assert list(sig.parameters) == ["basis_layout", "cutoff_fp64", "cutoff_fp32", "pair_wide_vk"]
    assert jk_pair.generate_get_j("L", 1e-7, 1e-13, pair_wide_vk=32)("mol", "dm", hermi=1)[0] == "J"
    assert jk_pair.generate_get_k("L", cutoff_fp64=1e-9)("mol", "dm")[1]["with_k"] is True
    assert seen == [("L", 1e-7, 1e-13), ("L", 1e-9, 1e-13)]


def test_shard_entry_partition(tmp_path):
    """The static multi-GPU partition (shard_entry, jqc_common.cuh) hands every task to exactly one rank."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "shard_entry_check")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-I", os.path.join(root, "joltqc_b200", "csrc"),
                    os.path.join(root, "tests", "shard_entry_check.cu"), "-o", exe], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr
