"""``jqc.pyscf.jk_pair`` — the reference's alternative, pair-list based entry points
(jqc/pyscf/jk_pair.py:49-115: ``generate_get_j / generate_get_k / generate_get_jk / generate_jk_kernel``
with the extra ``pair_wide_vk`` knob), same closure contract as ``jqc.pyscf.jk`` (jk_pair.py:117-135).

In the reference this is a second algorithm: J and K separately from q-sorted shell-pair lists with
register accumulation and a block reduction (jqc/backend/jk/pair_vj.cu, pair_vk.cu), which gives up the
8-fold symmetry and the per-quartet density screening to save atomics.  Here the pair-list design with
stationary outputs IS the default kernel path (``jk_brick.cuh`` / ``jk_bwarp.cuh``: shell-pair lists in ket
and bra order, lane-stationary J_kl / K_ik / K_il, warp-reduced J_ij) and keeps both the symmetry and the
screening, so these generators return the same engine-backed closures as ``joltqc_b200.pyscf.jk``.
``pair_wide_vk`` (the reference's K tile width) has no counterpart and is accepted and ignored.  Unlike the
reference's pair kernels (jqc/pyscf/tests/test_jk_pair.py:108, skipped there) several density matrices are
supported.
"""
from . import jk as _jk

__all__ = ["generate_get_j", "generate_get_k", "generate_get_jk", "generate_jk_kernel"]

PAIR_CUTOFF = _jk.PAIR_CUTOFF   # jk_pair.py:44
PAIR_WIDE_VK = 64               # jk_pair.py:46 (unused here)


def generate_jk_kernel(basis_layout, cutoff_fp64=1e-13, cutoff_fp32=1e-13, pair_wide_vk=PAIR_WIDE_VK):
    return _jk.generate_jk_kernel(basis_layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)


def generate_get_jk(basis_layout, cutoff_fp64=1e-13, cutoff_fp32=1e-13, pair_wide_vk=PAIR_WIDE_VK):
    return _jk.generate_get_jk(basis_layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)


def generate_get_j(basis_layout, cutoff_fp64=1e-13, cutoff_fp32=1e-13, pair_wide_vk=PAIR_WIDE_VK):
    return _jk.generate_get_j(basis_layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)


def generate_get_k(basis_layout, cutoff_fp64=1e-13, cutoff_fp32=1e-13, pair_wide_vk=PAIR_WIDE_VK):
    return _jk.generate_get_k(basis_layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)
