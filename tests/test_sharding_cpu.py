"""Multi-GPU path on CPU: world_size-2 gloo run of the partial-build -> all_reduce -> finalize
algebra with the oracle standing in for the per-rank engine (same static interleaved share)."""
import os
import socket

import numpy as np
import pytest

from tests.common import H2O, make, random_dm


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import OracleJK
    mol, lay = make(H2O, "def2-svp")
    orc = OracleJK(lay)
    dm = random_dm(mol.nao, 3)
    T = orc.transform()
    dmi = (T @ dm @ T.T)[None]
    # rank-local share: every world-th (ij) shell pair, like jqc_engine_set_shard's tile interleave
    vj, vk = orc.build_raw(dmi, 1, True, True, None, 1e-13, stride=world, phase=rank)
    buf = torch.from_numpy(np.concatenate([vj.ravel(), vk.ravel()]))
    dist.all_reduce(buf)                      # the one collective of the path
    n = vj.size
    vj = buf[:n].numpy().reshape(vj.shape)
    vk = buf[n:].numpy().reshape(vk.shape)
    J = T.T @ (2 * vj[0] + 2 * vj[0].T) @ T   # finalize (jk.py:353-370) after the reduction
    K = T.T @ (vk[0] + vk[0].T) @ T
    if rank == 0:
        np.save(out, np.stack([J, K]))
    dist.destroy_process_group()


def test_two_rank_partial_allreduce(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "jk.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    from oracle.oracle import OracleJK
    mol, lay = make(H2O, "def2-svp")
    rj, rk = OracleJK(lay).get_jk(random_dm(mol.nao, 3), 1)
    assert np.abs(got[0] - rj).max() < 1e-11 and np.abs(got[1] - rk).max() < 1e-11
