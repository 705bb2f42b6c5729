#!/usr/bin/env python
"""Golden J/K vectors for small cases, frozen from the CPU oracle (oracle/jk_oracle.c) after it was pinned against
the reference's golden RHF energies and Rys tables (tests/test_oracle.py) and, on the GPU, element-wise against the
reference's own kernels (tests/test_ref_kernels_gpu.py).  The reference itself cannot produce them here: it needs
CuPy + PySCF + a GPU.  Writes tests/golden/jk_small_cases.npz:
    <case>_dm, <case>_vj, <case>_vk   for the cases below (seeded dm = R R^T as jqc/pyscf/tests/test_jk.py:68-71).
usage: python tests/golden/make_jk_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleJK  # noqa: E402
from tests.common import H2_BOHR, H2O, make, random_dm  # noqa: E402

CASES = {
    # name: (atoms, basis, cart, unit, seed, hermi, omega)
    "h2_tzvpp_cart": (H2_BOHR, "def2-tzvpp", True, "B", 9, 1, None),       # jqc/pyscf/tests/test_jk.py:57-84
    "h2o_svp_sph": (H2O, "def2-svp", False, "Angstrom", 42, 1, None),
    "h2o_tzvpp_sph_lr": (H2O, "def2-tzvpp", False, "Angstrom", 7, 1, 0.5),  # long-range, omega = 0.5
}


def main():
    out = {}
    for name, (atom, basis, cart, unit, seed, hermi, omega) in CASES.items():
        mol, lay = make(atom, basis, cart=cart, unit=unit)
        dm = random_dm(mol.nao, seed)
        vj, vk = OracleJK(lay).get_jk(dm, hermi, True, True, omega, 1e-13)
        out[name + "_dm"], out[name + "_vj"], out[name + "_vk"] = dm, vj, vk
        print(name, dm.shape, float(np.abs(vj).max()), float(np.abs(vk).max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "jk_small_cases.npz"), **out)


if __name__ == "__main__":
    main()
