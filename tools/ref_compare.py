"""Reference JoltQC kernels vs the engine on one workload: time and element-wise parity.
usage: ref_compare.py [workload] [dm]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from joltqc_b200.pyscf.basis import BasisLayout
from oracle.ref_kernels import runner

wl = sys.argv[1] if len(sys.argv) > 1 else "valinomycin-tzvp"
mol, label = bench.build_mol(wl)
lay = BasisLayout.from_mol(mol)
eng = lay.engine()
dm = torch.as_tensor(bench.make_dm(mol, sys.argv[2] if len(sys.argv) > 2 else "ones"), device="cuda")
eng.q_matrix(0.0)
ref = runner.RefJK(lay, eng)
print("missing kernels:", len(ref.missing_kernels()))
dk = eng.dm_from_mol(dm)
t0 = time.perf_counter(); rj, rk = ref.get_jk_raw(dk, time_it=True); print("ref warm-up", ref.last, time.perf_counter() - t0)
rj, rk = ref.get_jk_raw(dk, time_it=True); print("ref timed", ref.last)
if len(sys.argv) > 3:
    ref.get_jk_raw(dk, time_classes=True)
    with open(sys.argv[3], "w") as f:
        f.write("class,ms,quartets\n")
        for cls, (ms, q) in sorted(ref.last["class_ms"].items(), key=lambda kv: -kv[1][0]):
            f.write("(%s|%s),%.3f,%d\n" % (cls[:2], cls[2:], ms, q))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    buf = eng.build_partial(dm, hermi=1)
    torch.cuda.synchronize(); t_ours = time.perf_counter() - t0
n2 = lay.nao ** 2
vj, vk = buf[:n2].reshape(lay.nao, lay.nao), buf[n2:2 * n2].reshape(lay.nao, lay.nao)
counts, _, _ = eng.last_stats()
out = {"workload": label, "ref_kernels_s": ref.last["seconds"], "ours_partial_s": t_ours, "ref_quartets": ref.last["quartets"],
       "ours_quartets": int(counts.sum()), "ref_launches": ref.last["launches"],
       "max_abs_dJ": (vj - rj).abs().max().item(), "max_abs_dK": (vk - rk).abs().max().item(),
       "max_abs_J": rj.abs().max().item(), "max_abs_K": rk.abs().max().item()}
print(json.dumps(out))
