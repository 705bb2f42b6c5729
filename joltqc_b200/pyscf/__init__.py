"""Plugin boundary: ``joltqc_b200.pyscf.apply(mf, config)``.

Same entry point and semantics as ``jqc.pyscf.apply`` (jqc/pyscf/__init__.py:121-254) for the
Coulomb/exchange path: it patches ``get_jk / get_j / get_k / get_veff / reset / as_scanner``
on a GPU4PySCF-style RHF/RKS object in place and returns it.  The DFT-grid kernels the
reference also patches on RKS objects (get_rho, nr_rks, nr_nlc_vxc) are out of scope here and
left to the host package; the RKS ``get_veff`` glue that decides which density reaches
``get_jk/get_j/get_k`` (jqc/pyscf/rks.py:180-262) is provided by ``joltqc_b200.pyscf.rks``.
"""
from functools import wraps
from types import MethodType
from typing import Any, Dict, Optional

__all__ = ["apply", "get_default_config"]


def get_default_config() -> Dict[str, Any]:
    """Same keys as the reference (jqc/pyscf/__init__.py:100-118) plus the multi-GPU knobs."""
    return {
        "jk": {
            "cutoff_fp32": None,  # None -> obj.direct_scf_tol
            "cutoff_fp64": None,
        },
        "dft": {"cutoff_fp32": 1e-13, "cutoff_fp64": 1e-6},   # accepted, unused (DFT kernels out of scope)
        "shard": None,            # (rank, world): evaluate one static share, all_reduce over torch.distributed
    }


def _wrap_reset(original_reset, config):
    @wraps(original_reset)
    def reset(self, mol=None):
        mf = original_reset(self, mol)
        return apply(mf, config)

    return reset


def _wrap_as_scanner(original_as_scanner, config):
    @wraps(original_as_scanner)
    def as_scanner(self, **kwargs):
        scanner = original_as_scanner(self, **kwargs)
        scanner._joltqc_applied = True
        if hasattr(scanner, "reset"):
            inner = scanner.reset.__func__ if hasattr(scanner.reset, "__func__") else scanner.reset
            scanner.reset = MethodType(_wrap_reset(inner, config), scanner)
        return scanner

    return as_scanner


def apply(obj, config: Optional[Dict[str, Any]] = None):
    """Patch ``obj`` in place with the B200 J/K engine and return it."""
    if "gpu4pyscf" not in obj.__class__.__module__ and hasattr(obj, "to_gpu"):
        obj = obj.to_gpu()
    if config is None:
        config = get_default_config()
    jk_cfg = config.get("jk", {}) or {}
    cutoff_fp32 = jk_cfg.get("cutoff_fp32")
    cutoff_fp64 = jk_cfg.get("cutoff_fp64")
    if cutoff_fp32 is None:
        cutoff_fp32 = getattr(obj, "direct_scf_tol", 1e-12)
    if cutoff_fp64 is None:
        cutoff_fp64 = getattr(obj, "direct_scf_tol", 1e-12)

    if hasattr(obj, "istype") and not obj.istype("RHF"):
        return obj

    from ..constants import TILE
    from . import jk as _jk
    from .basis import BasisLayout

    if hasattr(obj, "istype") and not obj.istype("DFRHF") and not obj.istype("DFRKS"):
        layout = BasisLayout.from_mol(obj.mol, alignment=TILE)
        shard = config.get("shard")
        if shard:
            layout.engine().enable_sharding(*shard)
        if hasattr(obj, "get_jk"):
            obj.get_jk = _jk.generate_jk_kernel(layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)
        if hasattr(obj, "get_j"):
            obj.get_j = _jk.generate_get_j(layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)
        if hasattr(obj, "get_k"):
            obj.get_k = _jk.generate_get_k(layout, cutoff_fp64=cutoff_fp64, cutoff_fp32=cutoff_fp32)
        if obj.istype("RKS"):
            from . import rks as _rks
            obj.get_veff = MethodType(_rks.generate_get_veff(), obj)
        elif obj.istype("RHF"):
            obj.get_veff = MethodType(_jk.generate_get_veff(), obj)
        obj._jqc_layout = layout

    obj._joltqc_applied = True
    if not hasattr(obj, "_jqc_original_reset"):
        original_reset = obj.reset.__func__
        obj._jqc_original_reset = original_reset
        obj.reset = MethodType(_wrap_reset(original_reset, config), obj)
    if hasattr(obj, "as_scanner"):
        orig = obj.as_scanner.__func__ if hasattr(obj.as_scanner, "__func__") else None
        if orig is not None and not getattr(orig, "_jqc_wrapped", False):
            wrapped = _wrap_as_scanner(orig, config)
            wrapped._jqc_wrapped = True
            obj.as_scanner = MethodType(wrapped, obj)
    return obj
