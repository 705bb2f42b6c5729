"""Run the UNMODIFIED reference J/K kernels (cubins in oracle/_ref/, built by build_ref_kernels.py
from the reference's own generators) on the local GPU, without CuPy.

Test/bench infrastructure, not product code: this is (i) the "reference JoltQC kernels on the same
B200" arm of bench.py and (ii) a GPU-side element-wise parity oracle at sizes the CPU oracle cannot
reach.  The host loop below restates the reference driver (jqc/pyscf/jk.py:162-348): density
pooling, make_tile_pairs (:385-431), the reversed group-quartet loop with 1024-tile chunks, one
screen_jk_tasks launch + BLOCKING read of `info` per chunk (:267-290), then the FP64 kernel on the
back of the 2 GiB queue (:315-328).  Launch shapes are the ones the reference closures use
(jk_1q1t.py:143-146, jk_1qnt.py:305-315, jk_tasks.py:85-106), recorded in manifest.json.
Kernels are loaded and launched through cuda-python (driver API) on torch-allocated buffers.
"""
import ctypes
import json
import math
import os

import numpy as np
import torch

try:
    from cuda.bindings import driver as cu
except ImportError:  # older cuda-python
    from cuda import cuda as cu

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref")
PAIR_CUTOFF = 1e-13          # jqc/pyscf/jk.py:48


def _ck(res):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError("CUDA driver error: %s" % cu.cuGetErrorString(err)[1])
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


def available():
    return os.path.exists(os.path.join(REF_DIR, "manifest.json"))


class RefKernels:
    def __init__(self):
        self.manifest = json.load(open(os.path.join(REF_DIR, "manifest.json")))
        self._fn = {}
        _ck(cu.cuInit(0))

    def has(self, name):
        ent = self.manifest["kernels"].get(name)
        return bool(ent) and os.path.exists(os.path.join(REF_DIR, ent["cubin"]))

    def function(self, name):
        if name not in self._fn:
            ent = self.manifest["kernels"][name]
            data = open(os.path.join(REF_DIR, ent["cubin"]), "rb").read()
            mod = _ck(cu.cuModuleLoadData(data))
            fn = _ck(cu.cuModuleGetFunction(mod, ent["entry"].encode()))
            if ent["max_dynamic_smem"] > 48 * 1024:
                _ck(cu.cuFuncSetAttribute(fn, cu.CUfunction_attribute.CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                          ent["max_dynamic_smem"]))
            self._fn[name] = (fn, ent)
        return self._fn[name]

    @staticmethod
    def launch(fn, grid, block, smem, stream, values, types):
        _ck(cu.cuLaunchKernel(fn, grid[0], grid[1] if len(grid) > 1 else 1, 1, block[0], block[1] if len(block) > 1 else 1, 1,
                              smem, stream, (tuple(values), tuple(types)), 0))


class RefJK:
    """get_jk of the reference on its own kernels, for one BasisLayout (kernel-side matrices)."""

    def __init__(self, layout, engine=None):
        self.lay = layout
        self.eng = engine or layout.engine()
        self.dev = self.eng.device
        self.k = RefKernels()
        self.nbas = int(layout.nbasis)
        self.nao = int(layout.nao)
        self.basis = torch.as_tensor(np.ascontiguousarray(layout.basis_data_fp64["packed"]), device=self.dev)
        self.gkey = [tuple(int(x) for x in k) for k in layout.group_key]
        self.goff = np.asarray(layout.group_offset, dtype=np.int64)
        loc = np.asarray(layout.ao_loc, dtype=np.int64)
        ao2shell = np.repeat(np.arange(self.nbas), np.diff(loc))
        self.ao2shell = torch.as_tensor(ao2shell, device=self.dev)
        self.queue_depth = int(self.k.manifest["queue_depth"])
        self.tile = int(self.k.manifest["tile"])
        self.chunk = int(self.k.manifest["max_pair_size"]) // (self.tile * self.tile)
        self.queue = None
        self.info = torch.zeros(4, dtype=torch.int32, device=self.dev)
        self.info_init = torch.tensor([0, 0, self.queue_depth, self.queue_depth], dtype=torch.int64).to(torch.int32)
        self.info_init = self.info_init.pin_memory()
        self.info_host = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.last = {}

    def missing_kernels(self):
        n = len(self.gkey)
        out = []
        for i in range(n):
            for j in range(i + 1):
                for k in range(i + 1):
                    for l in range(k + 1):
                        name = "jk_%d%d%d%d_%d%d%d%d" % (self.gkey[i][0], self.gkey[j][0], self.gkey[k][0], self.gkey[l][0],
                                                       self.gkey[i][1], self.gkey[j][1], self.gkey[k][1], self.gkey[l][1])
                        if not self.k.has(name):
                            out.append(name)
        return out

    # max_block_pooling + log (jk.py:172-184, linalg_helper.py:125-211)
    def _log_dm_cond(self, dms):
        a = dms.to(torch.float32).abs().amax(dim=0)
        nb = self.nbas
        rows = torch.zeros((nb, a.shape[1]), dtype=torch.float32, device=self.dev)
        rows.index_reduce_(0, self.ao2shell, a, "amax", include_self=True)
        cond = torch.zeros((nb, nb), dtype=torch.float32, device=self.dev)
        cond.index_reduce_(1, self.ao2shell, rows, "amax", include_self=True)
        return cond

    def _tile_pairs(self, q, cutoff):
        t = self.tile
        nt = self.nbas // t
        tq = q.reshape(nt, t, nt, t).amax(dim=(1, 3))
        tloc = self.goff // t
        pairs = {}
        for i in range(len(self.gkey)):
            for j in range(i + 1):
                sub = tq[tloc[i]:tloc[i + 1], tloc[j]:tloc[j + 1]]
                mask = sub > cutoff
                if i == j:
                    mask = torch.tril(mask)
                if not bool(mask.any()):
                    continue
                ii = torch.arange(tloc[i], tloc[i + 1], device=self.dev, dtype=torch.int32)
                jj = torch.arange(tloc[j], tloc[j + 1], device=self.dev, dtype=torch.int32)
                tij = ii[:, None] * nt + jj[None, :]
                order = torch.argsort(sub[mask])
                pairs[i, j] = tij[mask][order].contiguous()
        return pairs

    def get_jk_raw(self, dm_kern, hermi=1, cutoff=1e-13, time_it=False, time_classes=False):
        """dm_kern: (nao, nao) kernel-side density (one matrix, hermi = 1).  Returns the raw
        accumulators (vj, vk) of the kernels, i.e. before jk.py:353-370."""
        assert hermi == 1 and dm_kern.dim() == 2
        dev = self.dev
        stream = torch.cuda.current_stream().cuda_stream
        dms = dm_kern.reshape(1, self.nao, self.nao).contiguous()
        q = self.eng.q_matrix(0.0)
        if time_it:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        cond = self._log_dm_cond(dms)
        log_dm = torch.log(cond)                       # float32; log(0) = -inf for zero-width pads
        log_max = np.float32(log_dm.max().item())
        pair_cut = np.float32(math.log(PAIR_CUTOFF)) - log_max
        tile_pairs = self._tile_pairs(q, float(pair_cut))
        vj = torch.zeros_like(dms)
        vk = torch.zeros_like(dms)
        if self.queue is None:
            self.queue = torch.empty(self.queue_depth * 4, dtype=torch.int16, device=dev)      # ushort4 x QUEUE_DEPTH = 2 GiB
        log_cut = np.float32(math.log(cutoff))
        scr_fn, scr_ent = self.k.function("screen_jk_tasks_11")
        n = len(self.gkey)
        tasks = [(i, j, k, l) for i in range(n) for j in range(i + 1) for k in range(i + 1) for l in range(k + 1)]
        nquartets = 0
        launches = 0
        class_events = []
        i32, f32, f64, vp = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_void_p
        for (i, j, k, l) in tasks[::-1]:
            if (i, j) not in tile_pairs or (k, l) not in tile_pairs:
                continue
            tij, tkl = tile_pairs[i, j], tile_pairs[k, l]
            name = "jk_%d%d%d%d_%d%d%d%d" % (self.gkey[i][0], self.gkey[j][0], self.gkey[k][0], self.gkey[l][0],
                                           self.gkey[i][1], self.gkey[j][1], self.gkey[k][1], self.gkey[l][1])
            fn, ent = self.k.function(name)
            nsq = ent["block"][0]
            for a0 in range(0, tij.numel(), self.chunk):
                ta = tij[a0:a0 + self.chunk]
                for b0 in range(0, tkl.numel(), self.chunk):
                    tb = tkl[b0:b0 + self.chunk]
                    self.info.copy_(self.info_init, non_blocking=True)
                    grid = ((ta.numel() + 15) // 16, (tb.numel() + 15) // 16)
                    RefKernels.launch(scr_fn, grid, (16, 16), 0, stream,
                                      (self.queue.data_ptr(), self.info.data_ptr(), self.nbas, ta.data_ptr(), tb.data_ptr(),
                                       int(ta.numel()), int(tb.numel()), q.data_ptr(), log_dm.data_ptr(), float(log_cut),
                                       float(log_cut), float(log_max)),
                                      (vp, vp, i32, vp, vp, i32, i32, vp, vp, f32, f32, f32))
                    self.info_host.copy_(self.info)          # blocking D2H, as info.get() in the reference (jk.py:280)
                    torch.cuda.current_stream().synchronize()
                    offset = int(self.info_host[2].item()) & 0xFFFFFFFF
                    n64 = self.queue_depth - offset
                    launches += 1
                    if n64 > 0:
                        if time_classes:
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                        RefKernels.launch(fn, ((n64 + nsq - 1) // nsq,), tuple(ent["block"]), ent["shared_mem"], stream,
                                          (self.nao, self.basis.data_ptr(), dms.data_ptr(), vj.data_ptr(), vk.data_ptr(), 0.0,
                                           self.queue.data_ptr() + offset * 8, int(n64)),
                                          (i32, vp, vp, vp, vp, f64, vp, i32))
                        if time_classes:
                            e1.record()
                            class_events.append((name[3:7], e0, e1, n64))
                        launches += 1
                        nquartets += n64
        self.last = {"quartets": nquartets, "launches": launches}
        if time_classes:
            torch.cuda.synchronize()
            cm = {}
            for cls, e0, e1, nq in class_events:
                ms, q = cm.get(cls, (0.0, 0))
                cm[cls] = (ms + e0.elapsed_time(e1), q + nq)
            self.last["class_ms"] = cm
        if time_it:
            ev1.record()
            torch.cuda.synchronize()
            self.last["seconds"] = ev0.elapsed_time(ev1) * 1e-3
        return vj[0], vk[0]
