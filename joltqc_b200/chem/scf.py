"""A small RHF driver that stands in for ``gpu4pyscf.scf.RHF`` (absent from the images).

It exposes just what ``joltqc_b200.pyscf.apply`` patches on a GPU4PySCF mean-field object
(jqc/pyscf/__init__.py:121-254): ``get_jk / get_j / get_k / get_veff / reset / as_scanner``,
``direct_scf_tol``, ``istype`` and ``kernel()``.  The two-electron part has NO
implementation here: the stock ``get_jk`` raises until ``apply()`` wires in the CUDA engine
(there is no CPU fallback in the product; tests inject the oracle explicitly).
"""
import numpy as np

from . import int1e


def _to_numpy(a):
    if hasattr(a, "detach"):
        return a.detach().cpu().numpy()
    return np.asarray(a)


class RHF:
    def __init__(self, mol):
        self.mol = mol
        self.conv_tol = 1e-9
        self.conv_tol_grad = None
        self.max_cycle = 50
        self.direct_scf = True
        self.direct_scf_tol = 1e-13
        self.verbose = getattr(mol, "verbose", 0)
        self.diis_space = 8
        self.e_tot = None
        self.converged = False
        self.mo_energy = self.mo_coeff = self.mo_occ = None
        self.scf_summary = {}
        self._h1e = self._s1e = None

    # --- identity helpers mirroring pyscf.lib.StreamObject -----------------------------
    def istype(self, name):
        return name in ("RHF", "SCF")

    def to_gpu(self):
        return self

    # --- integrals ---------------------------------------------------------------------
    def _ensure_1e(self):
        if self._h1e is None:
            s, t, v = int1e.int1e(self.mol)
            self._s1e, self._h1e = s, t + v

    def get_ovlp(self, mol=None):
        self._ensure_1e()
        return self._s1e

    def get_hcore(self, mol=None):
        self._ensure_1e()
        return self._h1e

    def get_jk(self, mol=None, dm=None, hermi=1, vhfopt=None, with_j=True, with_k=True, omega=None):
        raise NotImplementedError(
            "joltqc_b200 has no CPU J/K path: call joltqc_b200.pyscf.apply(mf) on a machine with a B200"
        )

    def get_j(self, mol=None, dm=None, hermi=1, omega=None):
        return self.get_jk(mol, dm, hermi, with_k=False, omega=omega)[0]

    def get_k(self, mol=None, dm=None, hermi=1, omega=None):
        return self.get_jk(mol, dm, hermi, with_j=False, omega=omega)[1]

    def get_veff(self, mol=None, dm=None, dm_last=None, vhf_last=None, hermi=1):
        if dm is None:
            dm = self.make_rdm1()
        vj, vk = self.get_jk(mol, dm, hermi)
        return _to_numpy(vj) - 0.5 * _to_numpy(vk)

    # --- SCF ---------------------------------------------------------------------------
    def make_rdm1(self, mo_coeff=None, mo_occ=None):
        mo_coeff = self.mo_coeff if mo_coeff is None else mo_coeff
        mo_occ = self.mo_occ if mo_occ is None else mo_occ
        c = mo_coeff[:, mo_occ > 0]
        return (c * mo_occ[mo_occ > 0]) @ c.T

    def get_init_guess(self):
        h, s = self.get_hcore(), self.get_ovlp()
        e, c = self._eig(h, s)
        occ = np.zeros_like(e)
        occ[: self.mol.nelectron // 2] = 2.0
        return self.make_rdm1(c, occ)

    @staticmethod
    def _eig(f, s):
        w, u = np.linalg.eigh(s)
        x = u / np.sqrt(w)
        e, c = np.linalg.eigh(x.T @ f @ x)
        return e, x @ c

    def energy_elec(self, dm, h1e, vhf):
        return float(np.einsum("ij,ji->", h1e, dm) + 0.5 * np.einsum("ij,ji->", vhf, dm))

    def kernel(self, dm0=None):
        mol = self.mol
        h, s = self.get_hcore(), self.get_ovlp()
        dm = self.get_init_guess() if dm0 is None else np.asarray(dm0)
        nocc = mol.nelectron // 2
        vhf = _to_numpy(self.get_veff(mol, dm))
        e_nuc = mol.energy_nuc()
        e_tot = self.energy_elec(dm, h, vhf) + e_nuc
        errs, focks = [], []
        self.converged = False
        for cycle in range(self.max_cycle):
            f = h + vhf
            err = f @ dm @ s - s @ dm @ f
            focks.append(f)
            errs.append(err)
            if len(focks) > self.diis_space:
                focks.pop(0)
                errs.pop(0)
            if len(focks) > 1:
                n = len(focks)
                B = -np.ones((n + 1, n + 1))
                B[n, n] = 0
                for i in range(n):
                    for j in range(n):
                        B[i, j] = np.vdot(errs[i], errs[j])
                rhs = np.zeros(n + 1)
                rhs[n] = -1
                try:
                    c = np.linalg.solve(B, rhs)[:n]
                    f = sum(ci * fi for ci, fi in zip(c, focks))
                except np.linalg.LinAlgError:
                    pass
            e, cmo = self._eig(f, s)
            occ = np.zeros_like(e)
            occ[:nocc] = 2.0
            dm_last, vhf_last = dm, vhf
            dm = self.make_rdm1(cmo, occ)
            # incremental Fock build exactly as the reference's wrapper does it
            vhf = _to_numpy(self.get_veff(mol, dm, dm_last, vhf_last))
            e_last = e_tot
            e_tot = self.energy_elec(dm, h, vhf) + e_nuc
            self.mo_energy, self.mo_coeff, self.mo_occ = e, cmo, occ
            if self.verbose >= 4:
                print(f"cycle {cycle + 1:3d}  E = {e_tot:.12f}  dE = {e_tot - e_last:.3e}")
            if abs(e_tot - e_last) < self.conv_tol and np.linalg.norm(err) < np.sqrt(self.conv_tol) * 10:
                self.converged = True
                break
        self.e_tot = e_tot
        return e_tot

    def energy_tot(self):
        return self.e_tot

    def reset(self, mol=None):
        if mol is not None:
            self.mol = mol
        self._h1e = self._s1e = None
        self.e_tot = None
        self.converged = False
        return self

    def as_scanner(self):
        mf = self

        class _Scanner:
            def __init__(self):
                self.base = mf
                self.mol = mf.mol

            def reset(self, mol=None):
                self.base.reset(mol)
                if mol is not None:
                    self.mol = mol
                return self

            def __call__(self, mol_or_geom):
                mol = mol_or_geom if hasattr(mol_or_geom, "_bas") else self.mol.copy().set_geom_(mol_or_geom)
                self.reset(mol)
                return self.base.kernel()

        return _Scanner()
