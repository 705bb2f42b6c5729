"""One warm-up + one profiled J/K build of a bench workload; writes the per-class device-time table
(same columns as bench.py --class-profile).  For A/B runs of library variants (JQC_LIB_PATH=...) and
engine knobs (JQC_BWARP=..., ...).  usage: class_profile.py out.csv [workload] [dm] [cutoff_fp64]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from joltqc_b200.pyscf.basis import BasisLayout

out = sys.argv[1]
wl = sys.argv[2] if len(sys.argv) > 2 else "valinomycin-tzvp"
dmk = sys.argv[3] if len(sys.argv) > 3 else "ones"
c64 = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-13
mol, label = bench.build_mol(wl)
lay = BasisLayout.from_mol(mol, alignment=4)
eng = lay.engine()
dm = torch.as_tensor(bench.make_dm(mol, dmk), device="cuda")
eng.q_matrix(0.0)
ms = bench.time_builds(eng, dm, 1, 1) if c64 <= 1e-13 else None
eng.set_profiling(True)
vj, vk = eng.get_jk(dm, hermi=1, cutoff_fp64=c64, cutoff_fp32=1e-13)
class_ms = eng.last_class_ms()
counts, pw, _ = eng.last_stats()
(n64, n32), (ms64, ms32) = eng.last_band_stats()
with open(out, "w") as f:
    f.write("class,ms,quartets,alg_flops,tflops,frac_of_probe_peak\n")
    rows = []
    for key in np.nonzero(counts)[0]:
        fe, fd = bench.flops_per_class(int(key))
        fl = float(pw[key]) * fe + float(counts[key]) * fd
        rows.append((float(class_ms[key]), int(key), int(counts[key]), fl))
    for m, k, q, fl in sorted(rows, reverse=True):
        tf = fl / (m * 1e-3) / 1e12 if m > 0 else 0.0
        f.write("(%d%d|%d%d),%.4f,%d,%.4e,%.4f,%.4f\n" % (k // 125, k // 25 % 5, k // 5 % 5, k % 5, m, q, fl, tf, tf / 34.2))
print(label, "unprofiled build ms", ms, "sum class ms %.1f" % sum(r[0] for r in rows), "quartets fp64/fp32", n64, n32,
      "band ms %.1f / %.1f" % (ms64, ms32), "checksum", float(vj.sum()), float(vk.sum()))
