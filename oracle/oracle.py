"""Python face of the CPU oracle (test infrastructure — see the header of jk_oracle.c).

``OracleJK(layout).get_jk(dm, ...)`` restates the reference's ``get_jk`` closure
(jqc/pyscf/jk.py:109-382) on the CPU: AO transform in (basis.py:419-450), density pooling,
Schwarz + density screening, Rys ERIs with 8-fold symmetry, the six contractions, and the
post-processing / back-transform of jk.py:353-370.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# libcint's real-spherical coefficients (normalized="sp"), written out numerically here on
# purpose: the product derives the same matrices from the closed form
# (joltqc_b200/backend/cart2sph.py) and tests/test_cart2sph.py checks one against the other.
# Reference: jqc/backend/common/cart2sph.cu:22-100.  Rows = cartesian (lx desc, ly desc).
_C2S = {
    0: {(0, 0): 1.0},
    1: {(0, 0): 1.0, (1, 1): 1.0, (2, 2): 1.0},
    2: {(1, 0): 1.092548430592079070, (4, 1): 1.092548430592079070,
        (0, 2): -0.315391565252520002, (3, 2): -0.315391565252520002, (5, 2): 0.630783130505040012,
        (2, 3): 1.092548430592079070, (0, 4): 0.546274215296039535, (3, 4): -0.546274215296039535},
    3: {(1, 0): 1.770130769779930531, (6, 0): -0.590043589926643510,
        (4, 1): 2.890611442640554055,
        (1, 2): -0.457045799464465739, (6, 2): -0.457045799464465739, (8, 2): 1.828183197857862944,
        (2, 3): -1.119528997770346170, (7, 3): -1.119528997770346170, (9, 3): 0.746352665180230782,
        (0, 4): -0.457045799464465739, (3, 4): -0.457045799464465739, (5, 4): 1.828183197857862944,
        (2, 5): 1.445305721320277020, (7, 5): -1.445305721320277020,
        (0, 6): 0.590043589926643510, (3, 6): -1.770130769779930530},
    4: {(1, 0): 2.503342941796704538, (6, 0): -2.503342941796704530,
        (4, 1): 5.310392309339791593, (11, 1): -1.770130769779930530,
        (1, 2): -0.946174695757560014, (6, 2): -0.946174695757560014, (8, 2): 5.677048174545360108,
        (4, 3): -2.007139630671867500, (11, 3): -2.007139630671867500, (13, 3): 2.676186174229156671,
        (0, 4): 0.317356640745612911, (3, 4): 0.634713281491225822, (5, 4): -2.538853125964903290,
        (10, 4): 0.317356640745612911, (12, 4): -2.538853125964903290, (14, 4): 0.846284375321634430,
        (2, 5): -2.007139630671867500, (7, 5): -2.007139630671867500, (9, 5): 2.676186174229156671,
        (0, 6): -0.473087347878780002, (5, 6): 2.838524087272680054, (10, 6): 0.473087347878780009,
        (12, 6): -2.838524087272680050,
        (2, 7): 1.770130769779930531, (7, 7): -5.310392309339791590,
        (0, 8): 0.625835735449176134, (3, 8): -3.755014412695056800, (10, 8): 0.625835735449176134},
}


def c2s_matrix(l):
    m = np.zeros(((l + 1) * (l + 2) // 2, 2 * l + 1))
    for (c, s), v in _C2S[l].items():
        m[c, s] = v
    return m


def build(force=False):
    so = os.path.join(_HERE, "_build", "libjkoracle.so")
    src = os.path.join(_HERE, "jk_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp, ip, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float)
        L.oracle_rys_roots.argtypes = [ctypes.c_int, ctypes.c_double, dp]
        L.oracle_eri_block.argtypes = [dp, ip, ip] + [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_int, dp]
        L.oracle_q_cond.argtypes = [ctypes.c_int, dp, ip, ip, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int,
                                    ctypes.POINTER(dp), ctypes.c_double, fp]
        L.oracle_dm_cond.argtypes = [ctypes.c_int, ip, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, fp, fp]
        L.oracle_screen_ok.argtypes = [ctypes.c_int, fp, fp] + [ctypes.c_int] * 6 + [ctypes.c_float]
        L.oracle_build_jk.restype = ctypes.c_long
        L.oracle_build_jk.argtypes = [ctypes.c_int, dp, ip, ip, ip, fp, fp, ctypes.c_float, dp, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_float, dp, dp,
                                      ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _LIB = L
    return _LIB


def use_all_cores():
    """Pin the OpenMP thread count to the cores this process may run on (torchrun sets
    OMP_NUM_THREADS=1, which silently made the CPU baseline single-threaded in round 1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_num_threads(n)
    return lib().oracle_num_threads()


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def rys_roots(n, x):
    rw = np.zeros(2 * n)
    lib().oracle_rys_roots(n, float(x), _p(rw, ctypes.c_double))
    return rw[0::2].copy(), rw[1::2].copy()


class OracleJK:
    """CPU restatement bound to a shell table (a joltqc_b200 BasisLayout or anything with the
    same host fields: basis_data_fp64['packed'], angs, nprims, ao_loc, pad_id, mol_ao_offset,
    mol_nao, _mol.cart)."""

    def __init__(self, layout):
        self.layout = layout
        self.rec = np.ascontiguousarray(layout.basis_data_fp64["packed"], dtype=np.float64)
        self.angs = np.ascontiguousarray(layout.angs, dtype=np.int32)
        self.nprims = np.ascontiguousarray(layout.nprims, dtype=np.int32)
        self.ao_loc = np.ascontiguousarray(layout.ao_loc, dtype=np.int32)
        self.pad = np.ascontiguousarray(layout.pad_id, dtype=np.uint8)
        self.nbas = int(self.angs.size)
        self.nao = int(self.ao_loc[-1])
        self.cart = bool(layout._mol.cart)
        self._q = {}
        self._T = None

    # AO transform matrix: internal cart (rows) <- molecule AO (cols); basis.py:419-480
    def transform(self):
        if self._T is None:
            T = np.zeros((self.nao, self.layout.mol_nao))
            off = self.layout.mol_ao_offset
            for s in range(self.nbas):
                if self.pad[s]:
                    continue
                l = int(self.angs[s])
                c = np.eye((l + 1) * (l + 2) // 2) if self.cart else c2s_matrix(l)
                T[self.ao_loc[s] : self.ao_loc[s] + c.shape[0], off[s] : off[s] + c.shape[1]] = c
            self._T = T
        return self._T

    def q_matrix(self, omega=0.0):
        omega = 0.0 if omega is None else float(omega)
        if omega not in self._q:
            q = np.zeros((self.nbas, self.nbas), dtype=np.float32)
            mats = [np.ascontiguousarray(c2s_matrix(l)) for l in range(5)]
            arr = (ctypes.POINTER(ctypes.c_double) * 5)(*[_p(m, ctypes.c_double) for m in mats])
            lib().oracle_q_cond(self.nbas, _p(self.rec, ctypes.c_double), _p(self.angs, ctypes.c_int),
                                _p(self.nprims, ctypes.c_int), _p(self.pad, ctypes.c_uint8), int(self.cart),
                                arr, omega, _p(q, ctypes.c_float))
            self._q[omega] = q
        return self._q[omega]

    def eri_block(self, i, j, k, l, omega=0.0, sym=False):
        n = [(int(self.angs[s]) + 1) * (int(self.angs[s]) + 2) // 2 for s in (i, j, k, l)]
        out = np.zeros(n)
        lib().oracle_eri_block(_p(self.rec, ctypes.c_double), _p(self.angs, ctypes.c_int), _p(self.nprims, ctypes.c_int),
                               i, j, k, l, float(omega or 0.0), int(sym), _p(out, ctypes.c_double))
        return out

    def dm_cond(self, dms_int, hermi):
        dms_int = np.ascontiguousarray(dms_int, dtype=np.float64)
        out = np.zeros((self.nbas, self.nbas), dtype=np.float32)
        mx = ctypes.c_float(0)
        lib().oracle_dm_cond(self.nbas, _p(self.ao_loc, ctypes.c_int), self.nao, _p(dms_int, ctypes.c_double),
                             dms_int.shape[0], int(hermi), _p(out, ctypes.c_float), ctypes.byref(mx))
        return out, np.float32(mx.value)

    def build_raw(self, dms_int, hermi, with_j, with_k, omega, cutoff, stride=1, phase=0):
        """Kernel-side accumulation (before jk.py:353-370).  dms_int: (n, nao, nao) internal."""
        log_dm, log_max = self.dm_cond(dms_int, hermi)
        if hermi != 1:
            dms_int = np.concatenate([dms_int, dms_int.transpose(0, 2, 1)])   # jk.py:189-192
        dms_int = np.ascontiguousarray(dms_int)
        n = dms_int.shape[0]
        vj = np.zeros_like(dms_int) if with_j else np.zeros(1)
        vk = np.zeros_like(dms_int) if with_k else np.zeros(1)
        counts = np.zeros(625, dtype=np.int64)
        q = self.q_matrix(omega)
        nq = lib().oracle_build_jk(self.nbas, _p(self.rec, ctypes.c_double), _p(self.angs, ctypes.c_int),
                                   _p(self.nprims, ctypes.c_int), _p(self.ao_loc, ctypes.c_int),
                                   _p(q, ctypes.c_float), _p(log_dm, ctypes.c_float), ctypes.c_float(log_max),
                                   _p(dms_int, ctypes.c_double), n, int(with_j), int(with_k),
                                   float(omega or 0.0), ctypes.c_float(np.float32(math.log(cutoff))),
                                   _p(vj, ctypes.c_double), _p(vk, ctypes.c_double), stride, phase,
                                   _p(counts, ctypes.c_long))
        self.last_counts, self.last_nquartets = counts, nq
        return (vj if with_j else None), (vk if with_k else None)

    def get_jk(self, dm, hermi=0, with_j=True, with_k=True, omega=None, cutoff=1e-13):
        """Same contract as the reference closure (jk.py:109-118); numpy in, numpy out."""
        dm = np.asarray(dm, dtype=np.float64)
        T = self.transform()
        dms = dm.reshape(-1, dm.shape[-2], dm.shape[-1])
        dms_int = np.stack([T @ d @ T.T for d in dms])
        vj, vk = self.build_raw(dms_int, hermi, with_j, with_k, omega, cutoff)
        nd = dms.shape[0]

        def fin(v, is_j):
            if hermi == 1:
                v = v * 2.0 if is_j else v
                v = v + v.transpose(0, 2, 1)
            else:
                v = v[:nd] + v[nd:].transpose(0, 2, 1)
                if is_j:
                    v = v + v.transpose(0, 2, 1)
            return np.stack([T.T @ m @ T for m in v]).reshape(dm.shape)

        return (fin(vj, True) if with_j else 0), (fin(vk, False) if with_k else 0)
