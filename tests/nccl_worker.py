"""Worker of tests/test_nccl_gpu.py (one process per GPU, launched by torch.distributed.run):
sharded J/K build + ONE NCCL all_reduce + finalize, compared on rank 0 with the unsharded engine
and with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import benzene, make, random_dm  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mol, lay = make(benzene(), "def2-tzvp")
    dm = random_dm(mol.nao, 5) / mol.nao
    eng = lay.engine()
    eng.enable_sharding(rank, world)
    vj, vk = eng.get_jk(dm, hermi=1)                     # build_partial -> dist.all_reduce -> finalize
    counts = torch.tensor([float(eng.last_stats()[0].sum())], device=vj.device, dtype=torch.float64)
    dist.all_reduce(counts)
    ok = True
    if rank == 0:
        from joltqc_b200.backend.engine import JKEngine
        from oracle.oracle import OracleJK
        full = JKEngine(lay)
        fj, fk = full.get_jk(dm, hermi=1)
        nq_full = int(full.last_stats()[0].sum())
        orc = OracleJK(lay)
        rj, rk = orc.get_jk(dm, 1)
        e1 = max((vj - fj).abs().max().item() / fj.abs().max().item(), (vk - fk).abs().max().item() / fk.abs().max().item())
        e2 = max(np.abs(vj.cpu().numpy() - rj).max() / max(1.0, np.abs(rj).max()),
                 np.abs(vk.cpu().numpy() - rk).max() / max(1.0, np.abs(rk).max()))
        ok = e1 < 1e-11 and e2 < 1e-10 and int(counts.item()) == nq_full == int(orc.last_nquartets)
        print("NCCL_WORKER world=%d vs_1gpu=%.2e vs_oracle=%.2e quartets=%d/%d/%d %s"
              % (world, e1, e2, int(counts.item()), nq_full, int(orc.last_nquartets), "OK" if ok else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
