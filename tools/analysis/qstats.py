"""Offline statistics of the screened quartet structure (design aid, CPU only)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from joltqc_b200.pyscf.basis import BasisLayout
from oracle.oracle import OracleJK

wl = sys.argv[1] if len(sys.argv) > 1 else "valinomycin-tzvp"
mol, label = bench.build_mol(wl)
lay = BasisLayout.from_mol(mol, alignment=4)
orc = OracleJK(lay)
cache = f"/tmp/q_{wl}.npy"
if os.path.exists(cache):
    q = np.load(cache)
else:
    t0 = time.time(); q = orc.q_matrix(0.0); print("q in", time.time() - t0); np.save(cache, q)
print(label, "nbas", lay.nbasis, "groups", lay.group_key.tolist(), lay.group_offset.tolist())
np.save(f"/tmp/coords_{wl}.npy", lay.coords)
