"""Run N J/K builds of a bench workload (for ncu / quick timing).  usage: one_build.py [workload] [n] [dm] [cutoff_fp64]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from joltqc_b200.pyscf.basis import BasisLayout
wl = sys.argv[1] if len(sys.argv) > 1 else "valinomycin-tzvp"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mol, label = bench.build_mol(wl)
lay = BasisLayout.from_mol(mol)
eng = lay.engine()
dm = torch.as_tensor(bench.make_dm(mol, sys.argv[3] if len(sys.argv) > 3 else "ones"), device="cuda")
eng.q_matrix(0.0)
for i in range(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vj, vk = eng.get_jk(dm, hermi=1, cutoff_fp64=float(sys.argv[4]) if len(sys.argv) > 4 else 1e-13, cutoff_fp32=1e-13)
    torch.cuda.synchronize(); print("build", i, time.perf_counter() - t0, "s", float(vj.sum()), float(vk.sum()))
