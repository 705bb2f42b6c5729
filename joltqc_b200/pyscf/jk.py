"""J/K drivers with the reference's interface (jqc/pyscf/jk.py), backed by the CUDA engine.

``generate_jk_kernel(basis_layout, cutoff_fp64, cutoff_fp32)`` returns the closure
``get_jk(mol_ref, dm, hermi=0, vhfopt=None, with_j=True, with_k=True, omega=None, verbose=None)``
with the same contract as jqc/pyscf/jk.py:93-118: ``dm`` in the molecule's AO basis (2-D or
3-D, any array type), results on the device shaped like ``dm``, the un-requested one being
the int ``0`` (jk.py:194, 380); ``mol_ref``, ``vhfopt`` and ``verbose`` are accepted and
ignored (the layout is captured at construction); ``omega < 0`` asserts (jk.py:133-134).

Everything between those two points — AO transforms, density pooling, tile-pair screening,
task generation, the Rys kernels, symmetrisation — runs inside libjoltqc_b200.so through the
C ABI; this module holds no arithmetic.
"""
import time

import numpy as np

from ..constants import TILE
from .basis import BasisLayout

__all__ = ["generate_jk_kernel", "generate_get_j", "generate_get_k", "generate_get_jk", "generate_get_veff",
           "get_j", "get_jk"]

PAIR_CUTOFF = 1e-13  # jk.py:48 — applied inside the engine (tile lists)


def generate_jk_kernel(basis_layout: BasisLayout, cutoff_fp64=1e-13, cutoff_fp32=1e-13):
    assert basis_layout.alignment % TILE == 0, "the J/K layout must be padded to TILE (jqc/pyscf/__init__.py:189)"
    engine = basis_layout.engine()
    mol = basis_layout._mol

    def get_jk(mol_ref=None, dm=None, hermi=0, vhfopt=None, with_j=True, with_k=True, omega=None, verbose=None):
        assert with_j or with_k
        if omega is not None:
            assert omega >= 0.0, "short ranged J/K not supported"
        t0 = time.perf_counter()
        vj, vk = engine.get_jk(dm, hermi=hermi, with_j=with_j, with_k=with_k, omega=omega,
                               cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)
        if type(dm).__module__.split(".")[0] == "cupy":
            # GPU4PySCF hands in (and expects back) CuPy arrays: zero-copy views of the results
            import cupy as cp
            vj = cp.from_dlpack(vj) if with_j else 0
            vk = cp.from_dlpack(vk) if with_k else 0
        if getattr(mol, "verbose", 0) >= 5:
            import torch
            torch.cuda.synchronize()
            print(f"vj = {with_j} and vk = {with_k} take {time.perf_counter() - t0:.3f} sec")
        return vj, vk

    get_jk.engine = engine
    return get_jk


def generate_get_j(basis_layout, cutoff_fp64=1e-13, cutoff_fp32=1e-13):
    kern = generate_jk_kernel(basis_layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)

    def get_j(*args, **kwargs):
        return kern(*args, with_j=True, with_k=False, **kwargs)[0]

    return get_j


def generate_get_k(basis_layout, cutoff_fp64=1e-13, cutoff_fp32=1e-13):
    kern = generate_jk_kernel(basis_layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)

    def get_k(*args, **kwargs):
        return kern(*args, with_j=False, with_k=True, **kwargs)[1]

    return get_k


def generate_get_jk(basis_layout, cutoff_fp64=1e-13, cutoff_fp32=1e-13):
    return generate_jk_kernel(basis_layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)


def generate_get_veff():
    """RHF ``get_veff`` with the incremental (delta-density) Fock build of jk.py:78-90.  Works in the
    array namespace of ``dm`` (CuPy with GPU4PySCF, torch, or numpy -> torch), like the reference,
    so that a CuPy density goes straight to ``get_jk`` and a CuPy potential comes back."""
    from .rks import _asarray_like, _xp

    def get_veff(mf, mol=None, dm=None, dm_last=None, vhf_last=None, hermi=1):
        if dm is None:
            dm = mf.make_rdm1()
        if _xp(dm) is np:
            import torch
            dm = torch.as_tensor(np.asarray(dm))
        if dm_last is not None and not isinstance(dm_last, (int, float)) and mf.direct_scf:
            dm = dm - _asarray_like(dm, dm_last)
        vj, vk = mf.get_jk(mol, dm, hermi)
        vhf = vj - 0.5 * vk
        if vhf_last is not None and not isinstance(vhf_last, (int, float)):
            vhf = vhf + _asarray_like(vhf, vhf_last)
        return vhf

    return get_veff


# module-level conveniences mirroring jqc.pyscf.jk.__all__
def get_jk(mol, dm, hermi=0, with_j=True, with_k=True, omega=None, cutoff=1e-13):
    layout = BasisLayout.from_mol(mol, alignment=TILE)
    return generate_jk_kernel(layout, cutoff, cutoff)(mol, dm, hermi, with_j=with_j, with_k=with_k, omega=omega)


def get_j(mol, dm, hermi=0, omega=None, cutoff=1e-13):
    return get_jk(mol, dm, hermi, True, False, omega, cutoff)[0]
