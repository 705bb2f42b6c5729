"""RKS ``get_veff`` — the DFT caller of the J/K path (reference: jqc/pyscf/rks.py:180-262).

Only the Coulomb/exchange glue is in scope here: the exchange-correlation quadrature
(``ks._numint.nr_rks`` / ``nr_nlc_vxc``, grid set-up) stays with the host package
(GPU4PySCF); the reference's own JIT kernels for those are out of scope.
What this wrapper decides is WHICH density reaches ``get_jk / get_j / get_k``:
the incremental delta-density build of direct SCF, pure vs hybrid functionals, and the extra
long-range exchange build of range-separated hybrids (``omega`` -> erf-attenuated kernels).
"""
import numpy as np

__all__ = ["generate_get_veff", "TaggedArray"]


def _xp(a):
    """array namespace helpers that work for torch tensors, CuPy and numpy arrays"""
    mod = type(a).__module__.split(".")[0]
    if mod == "torch":
        import torch
        return torch
    if mod == "cupy":
        import cupy
        return cupy
    return np


def _asarray_like(ref, a):
    if isinstance(a, (int, float)):
        return a
    xp = _xp(ref)
    if xp.__name__ == "torch":
        import torch
        return a.to(ref.device) if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a), device=ref.device)
    return xp.asarray(a)


def _trace_prod(a, b):
    xp = _xp(a)
    if xp.__name__ == "torch":
        return float((a * b.transpose(-1, -2)).sum().real)
    return float(xp.einsum("ij,ji", a, b).real)


class TaggedArray:
    """Minimal stand-in for gpu4pyscf.lib.cupy_helper.tag_array: the effective potential plus the
    attributes the SCF driver reads back (ecoul, exc, vj, vk).  With GPU4PySCF installed the real
    ``tag_array`` is used instead."""

    def __init__(self, array, **attrs):
        self.array = array
        self.__dict__.update(attrs)

    def __array__(self, dtype=None):
        a = self.array
        a = a.detach().cpu().numpy() if hasattr(a, "detach") else (a.get() if hasattr(a, "get") else np.asarray(a))
        return a.astype(dtype) if dtype else a

    def __getattr__(self, name):            # delegate shape, dtype, ... to the wrapped array
        return getattr(self.__dict__["array"], name)

    def __add__(self, other):
        return self.array + (other.array if isinstance(other, TaggedArray) else other)

    __radd__ = __add__


def _tag(vxc, **attrs):
    try:
        from gpu4pyscf.lib.cupy_helper import tag_array
        if type(vxc).__module__.split(".")[0] == "cupy":
            return tag_array(vxc, **attrs)
    except ImportError:
        pass
    return TaggedArray(vxc, **attrs)


def generate_get_veff():
    def get_veff(ks, mol=None, dm=None, dm_last=0, vhf_last=0, hermi=1):
        if mol is None:
            mol = ks.mol
        if dm is None:
            dm = ks.make_rdm1()
        try:                                   # grid set-up belongs to the host package
            from gpu4pyscf.dft.rks import initialize_grids
            initialize_grids(ks, mol, dm)
        except ImportError:
            if getattr(ks, "grids", None) is not None and getattr(ks.grids, "coords", None) is None and hasattr(ks.grids, "build"):
                ks.grids.build()
        ground_state = getattr(dm, "ndim", 0) == 2
        ni = ks._numint
        if hermi == 2:                         # because rho = 0
            n, exc, vxc = 0, 0, 0
        else:
            n, exc, vxc = ni.nr_rks(mol, ks.grids, ks.xc, dm)
            if hasattr(ks, "do_nlc") and ks.do_nlc():
                xc = ks.xc if ni.libxc.is_nlc(ks.xc) else ks.nlc
                n, enlc, vnlc = ni.nr_nlc_vxc(mol, ks.nlcgrids, xc, dm)
                exc = exc + enlc
                vxc = vxc + vnlc
        incremental = getattr(ks, "_eri", None) is None and getattr(ks, "direct_scf", True)
        if not ni.libxc.is_hybrid_xc(ks.xc):
            vk = None
            if incremental and getattr(vhf_last, "vj", None) is not None:
                ddm = _asarray_like(vhf_last.vj, dm) - _asarray_like(vhf_last.vj, dm_last)
                vj = ks.get_j(mol, ddm, hermi)
                vj = vj + vhf_last.vj
            else:
                vj = ks.get_j(mol, dm, hermi)
            vxc = _asarray_like(vj, vxc) + vj
        else:
            omega, alpha, hyb = ni.rsh_and_hybrid_coeff(ks.xc, spin=getattr(mol, "spin", 0))
            if incremental and getattr(vhf_last, "vk", None) is not None:
                ddm = _asarray_like(vhf_last.vk, dm) - _asarray_like(vhf_last.vk, dm_last)
                vj, vk = ks.get_jk(mol, ddm, hermi)
                vk = vk * hyb
                if abs(omega) > 1e-10:         # range-separated Coulomb operator
                    vk = vk + ks.get_k(mol, ddm, hermi, omega=omega) * (alpha - hyb)
                vj = vj + vhf_last.vj
                vk = vk + vhf_last.vk
            else:
                vj, vk = ks.get_jk(mol, dm, hermi)
                vk = vk * hyb
                if abs(omega) > 1e-10:
                    vk = vk + ks.get_k(mol, dm, hermi, omega=omega) * (alpha - hyb)
            vxc = _asarray_like(vj, vxc) + vj - vk * 0.5
            if ground_state:
                exc = exc - _trace_prod(_asarray_like(vk, dm), vk) * 0.5 * 0.5
        ecoul = _trace_prod(_asarray_like(vj, dm), vj) * 0.5 if ground_state else None
        return _tag(vxc, ecoul=ecoul, exc=exc, vj=vj, vk=vk)

    return get_veff
