"""Compile the UNMODIFIED reference J/K kernels (JoltQC) for sm_100a into oracle/_ref/.

Test/bench infrastructure, not product code.  Runs only where /root/reference exists (the build
container): the reference's own Python generators (jqc/backend/jk.py -> jk_1q1t.py / jk_1qnt.py,
jk_tasks.py, util.py) are imported unchanged with a stub `cupy` module that records the CUDA text
handed to cp.RawModule and the launch geometry of the returned closures, i.e. the routing
(A100 FP64 fragment table for an unknown device, jqc/backend/jk.py:46-50), the generated constexpr
header (jk_1q1t.py:51-77, jk_1qnt.py:237-276) and the kernel text (jqc/backend/jk/*.cu,
jqc/backend/rys/*.cu) are exactly what the reference would JIT on a B200.  The text is compiled
with nvcc (-std=c++17 --use_fast_math, the reference's NVRTC options) to one cubin per
(angular momenta, primitive counts) key.  No reference source is copied into the repository;
oracle/_ref/ is git-ignored and travels to the GPU box with the snapshot.

usage: python -m oracle.ref_kernels.build_ref_kernels [workload ...]
"""
import hashlib
import json
import os
import subprocess
import sys
import types
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "oracle", "_ref")
REF = os.environ.get("JQC_REFERENCE", "/root/reference")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

_captured = []     # (code, options) in RawModule creation order


class _Kernel:
    def __init__(self, mod, name):
        self.mod, self.name = mod, name
        self.local_size_bytes = 0
        self.num_regs = 0
        self.max_dynamic_shared_size_bytes = 0
        self.launches = []

    def __call__(self, grid, block, args, shared_mem=0):
        self.launches.append((tuple(grid), tuple(block), int(shared_mem)))


class _RawModule:
    def __init__(self, code=None, options=(), **kw):
        self.code, self.options, self.kernels = code, tuple(options), {}
        _captured.append(self)

    def get_function(self, name):
        self.kernels.setdefault(name, _Kernel(self, name))
        return self.kernels[name]


class _Any:
    """Permissive stand-in for everything else the reference touches on cupy at import time."""

    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


def _install_stubs():
    cp = types.ModuleType("cupy")
    cp.RawModule = _RawModule
    cuda = types.SimpleNamespace()
    cuda.Device = lambda *a: types.SimpleNamespace(id=0)
    cuda.runtime = types.SimpleNamespace(
        getDevice=lambda: 0,
        # a device name the reference ships no table for -> it falls back to the A100 FP64 table,
        # which is what happens on a real B200; 48 KB is sharedMemPerBlock of every CUDA device
        getDeviceProperties=lambda i: {"name": b"NVIDIA B200", "sharedMemPerBlock": 48 * 1024})
    cuda.alloc_pinned_memory = lambda n: bytearray(n)
    cp.cuda = cuda
    cp.__getattr__ = lambda k: _Any()
    sys.modules["cupy"] = cp
    # package shells so that jqc/__init__.py and jqc/backend/__init__.py (which pull in DFT/ECP
    # modules) are not executed; submodules are imported from the reference tree unchanged
    for name, sub in (("jqc", "jqc"), ("jqc.backend", "jqc/backend")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m


def group_keys(workload):
    sys.path.insert(0, ROOT)
    import bench
    from joltqc_b200.pyscf.basis import BasisLayout
    mol, _ = bench.build_mol(workload)
    lay = BasisLayout.from_mol(mol, alignment=4)
    gk = [tuple(int(x) for x in k) for k in lay.group_key]
    n = len(gk)
    keys = set()
    for i in range(n):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    keys.add(((gk[i][0], gk[j][0], gk[k][0], gk[l][0]), (gk[i][1], gk[j][1], gk[k][1], gk[l][1])))
    return sorted(keys)


def key_name(ang, nprim):
    return "jk_%d%d%d%d_%d%d%d%d" % (*ang, *nprim)


def _nvcc(job):
    src, cubin = job
    r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--use_fast_math", "-cubin",
                        "-o", cubin, src], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stderr[-2000:]))
    return cubin


def main(workloads):
    if not os.path.isdir(REF):
        raise SystemExit("the reference tree is not present: nothing to build (the prebuilt oracle/_ref is used)")
    _install_stubs()
    from jqc.backend.jk import gen_jk_kernel          # reference router, unchanged
    from jqc.backend.jk_tasks import gen_screen_jk_tasks_kernel, QUEUE_DEPTH, MAX_PAIR_SIZE
    from jqc.constants import TILE
    os.makedirs(os.path.join(OUT, "src"), exist_ok=True)
    mpath = os.path.join(OUT, "manifest.json")
    manifest = json.load(open(mpath)) if os.path.exists(mpath) else {"kernels": {}}
    manifest.update({"queue_depth": int(QUEUE_DEPTH), "max_pair_size": int(MAX_PAIR_SIZE), "tile": int(TILE),
                     "compile_options": ["-std=c++17", "--use_fast_math"], "reference": REF})
    jobs = []

    def emit(name, mod, kern, launch):
        code = mod.code
        h = hashlib.sha1(code.encode()).hexdigest()[:16]
        src = os.path.join(OUT, "src", name + ".cu")
        cubin = os.path.join(OUT, name + ".cubin")
        ent = manifest["kernels"].get(name)
        if not (ent and ent.get("sha") == h and os.path.exists(cubin)):
            with open(src, "w") as f:
                f.write(code)
            jobs.append((src, cubin))
        grid, block, smem = launch
        manifest["kernels"][name] = {"sha": h, "cubin": name + ".cubin", "entry": kern.name, "block": list(block),
                                     "shared_mem": smem, "max_dynamic_smem": int(kern.max_dynamic_shared_size_bytes)}

    # task generator (do_j = do_k = 1, omega = None), launched once with dummy arguments to record the geometry
    n0 = len(_captured)
    _, _, fun = gen_screen_jk_tasks_kernel(do_j=True, do_k=True, tile=TILE)
    mod = _captured[n0]
    kern = mod.kernels["screen_jk_tasks"]
    manifest["kernels"]["screen_jk_tasks_11"] = None
    emit("screen_jk_tasks_11", mod, kern, ((1, 1), (16, 16), 0))
    for wl in workloads:
        for ang, nprim in group_keys(wl):
            name = key_name(ang, nprim)
            n0 = len(_captured)
            fun = gen_jk_kernel(ang, nprim, dtype=np.float64, n_dm=1, do_j=True, do_k=True, omega=None)
            if len(_captured) == n0:
                continue                                 # lru_cache hit: already emitted in this run
            mod = _captured[n0]
            kern = next(iter(mod.kernels.values()))
            fun(0, 0, 0, 0, 0, 0, 0, 1)                  # records (grid, block, shared_mem); last arg = ntasks
            emit(name, mod, kern, kern.launches[-1])
    print("compiling %d kernels with nvcc ..." % len(jobs), flush=True)
    with ThreadPoolExecutor(int(os.environ.get("JOBS", os.cpu_count()))) as ex:
        for i, c in enumerate(ex.map(_nvcc, jobs)):
            if i % 20 == 0:
                print("  %d/%d %s" % (i + 1, len(jobs), os.path.basename(c)), flush=True)
    # the CUDA text is an intermediate: only cubins + manifest stay (no reference source in the tree)
    for src, _ in jobs:
        os.remove(src)
    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    print("oracle/_ref: %d kernels" % len(manifest["kernels"]))


if __name__ == "__main__":
    main(sys.argv[1:] or ["taxol-svp", "valinomycin-tzvp"])
