#!/usr/bin/env python
"""Benchmark of the FP64 direct-SCF J/K build (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference|jqc-kernels] [--workload NAME] [--dm ones|decay]

A "step" is one get_jk (J and K, hermi=1, cutoff 1e-13) of the workload.  Default workload:
config 4 of BASELINE.json — valinomycin/def2-TZVP, represented by the in-tree stand-in
geometry C62H86N2O16 (166 atoms, nao 2996; the real geometry is not available offline).
D is synthetic: all ones, exactly what the reference's own J/K benchmark feeds
(benchmarks/benchmark_jk.py:99-132), or an exponentially decaying model density (--dm decay).

Printed JSON (one line, rank 0): see the measurement contract in the task statement; `value`
is seconds per J/K build with D resident in HBM (device-timed, max over ranks), `e2e` the
same through the host-buffer C-ABI call (jqc_get_jk_host: pinned host D in, J and K out),
`roofline` the algorithmic FP64 FLOP/s of the build (SURVEY 8d model) against the FP64 FMA
peak measured on the box, `cpu_baseline` the CPU restatement (oracle) on the host cores
extrapolated from a bounded sample (with one COMPLETE small build as the anchor of the
extrapolation), `ref_kernels` the UNMODIFIED JoltQC kernels (oracle/_ref cubins) on the same GPU
and the element-wise difference of the two results, `extra` the other BASELINE.json configs.

--impl reference   : the CPU port alone (the reference's CPU path, libcint, is not installed).
--impl jqc-kernels : the unmodified reference kernels alone (1 GPU).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (xyz file or builder, basis, label)
    "valinomycin-tzvp": ("valinomycin_standin_C62H86N2O16.xyz", "def2-tzvp",
                         "valinomycin/def2-TZVP J+K (stand-in geometry C62H86N2O16, 166 atoms)"),
    "valinomycin-tzvpp": ("valinomycin_standin_C62H86N2O16.xyz", "def2-tzvpp",
                          "valinomycin/def2-TZVPP J+K (stand-in geometry C62H86N2O16)"),
    "taxol-svp": ("taxol_standin_C46H52N8O6.xyz", "def2-svp", "taxol/def2-SVP J+K (stand-in geometry C46H52N8O6, 112 atoms)"),
    "benzene-ccpvtz": ("benzene", "cc-pvtz", "benzene/cc-pVTZ J+K"),
    "h2o-tzvpp": ("h2o", "def2-tzvpp", "H2O/def2-TZVPP J+K"),
}


def nf(l):
    return (l + 1) * (l + 2) // 2


def flops_per_class(key, do_j=True, do_k=True, n_dm=1):
    """(per primitive quartet ERI flops, per quartet digestion flops) — SURVEY.md 8(d)."""
    ll = key % 5; lk = key // 5 % 5; lj = key // 25 % 5; li = key // 125
    N = nf(li) * nf(lj) * nf(lk) * nf(ll)
    L = li + lj + lk + ll
    nroots = L // 2 + 1
    lij, lkl = li + lj, lk + ll
    R = 76 * nroots
    t_ij = 0 if lij == 0 else 1 + 4 * (lij - 1)
    t_kl = 0 if lkl == 0 else (1 + 4 * (lkl - 1)) + lij * (4 + 6 * (lkl - 1))
    T = 3 * (t_ij + t_kl) + 20
    h_j = (lkl + 1) * sum(lij - j for j in range(lj))
    h_l = (li + 1) * (lj + 1) * sum(lkl - l for l in range(ll))
    H = 6 * (h_j + h_l)
    return R + nroots * (T + H + 3 * N), n_dm * N * (4 * do_j + 8 * do_k)


def total_flops(counts, prim_weighted, do_j=True, do_k=True, n_dm=1):
    tot = 0.0
    for key in np.nonzero(counts)[0]:
        fe, fd = flops_per_class(int(key), do_j, do_k, n_dm)
        tot += float(prim_weighted[key]) * fe + float(counts[key]) * fd
    return tot


def build_mol(name):
    from joltqc_b200.chem.mole import M, read_xyz
    src, basis, label = WORKLOADS[name]
    if src == "benzene":
        rc, rh = 1.39, 2.48
        atom = [("C", (rc * math.cos(math.pi / 3 * k), rc * math.sin(math.pi / 3 * k), 0.0)) for k in range(6)]
        atom += [("H", (rh * math.cos(math.pi / 3 * k), rh * math.sin(math.pi / 3 * k), 0.0)) for k in range(6)]
    elif src == "h2o":
        atom = "O 0 0 0.1174; H -0.757 0 -0.4696; H 0.757 0 -0.4696"
    else:
        atom = read_xyz(os.path.join(ROOT, "joltqc_b200", "chem", "molecules", src))
    return M(atom=atom, basis=basis), label


def make_dm(mol, kind):
    nao = mol.nao
    if kind == "ones":
        return np.ones((nao, nao))
    # decaying model density: |D_mu,nu| ~ exp(-0.6 |R_A - R_B|), seeded signs, strong diagonal
    rng = np.random.RandomState(7)
    loc = mol.ao_loc
    atom_of = np.repeat(mol._bas[:, 0], np.diff(loc))
    xyz = mol.atom_coords()[atom_of]
    dist = np.linalg.norm(xyz[:, None, :] - xyz[None, :, :], axis=-1)
    r = rng.randn(nao, nao)
    dm = 0.3 * (r + r.T) * np.exp(-0.6 * dist)
    dm[np.arange(nao), np.arange(nao)] += 1.0
    return dm


class ClockSampler:
    """nvidia-smi clocks/throttle sampling during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


def cpu_port_baseline(lay, dm, budget_s=20.0, stride=None):
    """Time the CPU restatement on a bounded sample (every stride-th (ij) pair) and extrapolate."""
    from oracle.oracle import OracleJK, use_all_cores
    cores = use_all_cores()
    orc = OracleJK(lay)
    T = orc.transform()
    dmi = (T @ dm @ T.T)[None]
    t0 = time.perf_counter()
    orc.q_matrix(0.0)
    t_q = time.perf_counter() - t0
    npair = lay.nbasis * (lay.nbasis + 1) // 2
    fixed = stride is not None
    stride = stride or max(1, npair // 400)
    while True:
        t0 = time.perf_counter()
        orc.build_raw(dmi, 1, True, True, None, 1e-13, stride=stride, phase=stride // 2)
        dt = time.perf_counter() - t0
        nq = int(orc.last_nquartets)
        if fixed or dt > budget_s / 4 or stride == 1:
            break
        stride = max(1, int(stride / min(8.0, max(2.0, budget_s / 2 / max(dt, 1e-3)))))
    return {"seconds_sample": dt, "quartets_sample": nq, "stride": stride, "cores": cores, "schwarz_seconds": t_q,
            "counts": orc.last_counts.copy()}


def cpu_anchor():
    """One COMPLETE (un-extrapolated) CPU build of benzene/cc-pVTZ next to its own sampled
    extrapolation: how far the stride sampling used for the big workloads is off."""
    from joltqc_b200.pyscf.basis import BasisLayout
    mol, label = build_mol("benzene-ccpvtz")
    lay = BasisLayout.from_mol(mol, alignment=4)
    dm = make_dm(mol, "ones")
    full = cpu_port_baseline(lay, dm, stride=1)
    samp = cpu_port_baseline(lay, dm, stride=16)
    return {"workload": label, "complete_build_s": full["seconds_sample"], "quartets": full["quartets_sample"],
            "sampled_stride": samp["stride"], "extrapolated_s": samp["seconds_sample"] * samp["stride"],
            "extrapolation_ratio": samp["seconds_sample"] * samp["stride"] / max(full["seconds_sample"], 1e-9)}


def ref_kernels_leg(lay, eng, dm_dev, reps=1):
    """Unmodified JoltQC kernels (oracle/_ref) on this GPU: seconds per build + element-wise
    comparison of the kernel-side J/K accumulators with the engine's."""
    import torch
    try:
        from oracle.ref_kernels import runner
    except Exception as exc:           # cuda-python missing
        return {"unavailable": "cuda-python runner not importable: %s" % exc}
    if not runner.available():
        return {"unavailable": "oracle/_ref not built (python -m oracle.ref_kernels.build_ref_kernels needs the reference tree)"}
    ref = runner.RefJK(lay, eng)
    missing = ref.missing_kernels()
    if missing:
        return {"unavailable": "oracle/_ref lacks %d kernels for this basis" % len(missing)}
    dk = eng.dm_from_mol(dm_dev)
    ref.get_jk_raw(dk)                                   # warm-up: loads the cubins
    ts = []
    for _ in range(reps):
        rj, rk = ref.get_jk_raw(dk, time_it=True)
        ts.append(ref.last["seconds"])
    buf = eng.build_partial(dm_dev, hermi=1)
    torch.cuda.synchronize()
    n2 = lay.nao ** 2
    vj, vk = buf[:n2].reshape(lay.nao, lay.nao), buf[n2:2 * n2].reshape(lay.nao, lay.nao)
    counts, _, _ = eng.last_stats()
    out = {"value": float(np.median(ts)), "unit": "s/iter", "kind": "unmodified JoltQC kernels (rys_1q1t_vjk / rys_1qnt_vjk / "
           "screen_jk_tasks, A100 FP64 scheme table, reference driver loop with its blocking info reads), nvcc sm_100a "
           "-std=c++17 --use_fast_math, launched through cuda-python",
           "quartets": int(ref.last["quartets"]), "same_quartet_count": bool(int(counts.sum()) == ref.last["quartets"]),
           "launches": int(ref.last["launches"]),
           "max_abs_dJ_kernel_side": float((vj - rj).abs().max().item()), "max_abs_dK_kernel_side": float((vk - rk).abs().max().item()),
           "max_abs_J": float(rj.abs().max().item()), "max_abs_K": float(rk.abs().max().item())}
    del ref
    torch.cuda.empty_cache()
    return out


_REAL_STDOUT = None


def _capture_stdout():
    """Route fd 1 to stderr for the whole run (NCCL and other native libraries print banners on
    stdout) and keep the original stdout for the single JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def _config(args, mol, lay, label, world):
    return {"workload": label, "basis": WORKLOADS[args.workload][1], "nao": mol.nao, "nao_cart_kernel": lay.nao,
            "shells": int((~lay.pad_id).sum()), "dm": args.dm, "hermi": 1, "with_j": True, "with_k": True,
            "cutoff": 1e-13, "parallelism": f"static task interleave over {max(world, 1)} GPU(s) + 1 NCCL all_reduce",
            "l2": "working set D+J+K = %.0f MB %s 126 MB L2, no explicit flush" % (3 * lay.nao**2 * 8 / 1e6, ">" if 3 * lay.nao**2 * 8 > 126e6 else "<")}


def run_reference_arm(args, lay, dm, config):
    """CPU port on all host cores: W untimed + K timed bounded samples of the same build."""
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    first = cpu_port_baseline(lay, dm, budget_s=per_step)          # chooses the stride for the budget
    for _ in range(max(0, args.warmup - 1)):
        cpu_port_baseline(lay, dm, stride=first["stride"])
    ts, res = [], first
    for _ in range(max(1, args.steps)):
        res = cpu_port_baseline(lay, dm, stride=first["stride"])
        ts.append(res["seconds_sample"])
    est = float(np.mean(ts)) * res["stride"]
    sample = (f"every {res['stride']}th (ij) shell pair of the same build ({res['quartets_sample']} quartets in "
              f"{np.mean(ts):.2f} s per step), extrapolated x{res['stride']}")
    _emit({"impl": "reference", "metric": "fp64_jk_build_time", "value": est, "unit": "s/iter", "n_gpus": 0,
           "steps": len(ts), "warmup": args.warmup, "ms_per_step": est * 1e3, "higher_is_better": False,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "cpu_baseline": {"value": est, "unit": "s/iter", "cores": res["cores"], "kind": "port", "sample": sample},
           "e2e": {"value": est, "unit": "s/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def time_builds(eng, dm_dev, warmup, steps):
    import torch
    for _ in range(warmup):
        eng.get_jk(dm_dev, hermi=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.get_jk(dm_dev, hermi=1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def mixed_leg(eng, dm_dev, vj64, vk64, cutoff_fp64=1e-7):
    """BASELINE config 5 column "FP64 vs mixed": the same build with the FP32 band of the reference
    (cutoff_fp32 = 1e-13 < estimate <= cutoff_fp64, jk.py:93-96) evaluated in single precision.
    Reported separately from the FP64 headline, with its element-wise deviation from the FP64 result."""
    import torch
    eng.get_jk(dm_dev, hermi=1, cutoff_fp64=cutoff_fp64, cutoff_fp32=1e-13)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vj, vk = eng.get_jk(dm_dev, hermi=1, cutoff_fp64=cutoff_fp64, cutoff_fp32=1e-13)
    e1.record()
    torch.cuda.synchronize()
    (n64, n32), _ = eng.last_band_stats()
    return {"s_per_build": e0.elapsed_time(e1) * 1e-3, "cutoff_fp64": cutoff_fp64, "cutoff_fp32": 1e-13,
            "quartets_fp64": n64, "quartets_fp32": n32,
            "fp32_kernels": "every angular class up to f shells (float integrals, FP64 accumulation: brick kernel <= 108 integrals, brick-scheduled multi-lane kernel above); classes with g shells and > 108 integrals evaluate their band in FP64",
            "max_abs_dJ_vs_fp64": float((vj - vj64).abs().max().item()), "max_abs_dK_vs_fp64": float((vk - vk64).abs().max().item()),
            "max_abs_J": float(vj64.abs().max().item()), "max_abs_K": float(vk64.abs().max().item()),
            "stated_tolerance": "max-abs 1e-7 x max(1, max|J|) on J and K (reference tests: 1e-7, jqc/pyscf/tests/test_jk.py:245-246)"}


def extra_configs(peak_probe, dev):
    """The other BASELINE.json configurations, 1 warm-up + 1 timed build each (1 GPU)."""
    import torch
    from joltqc_b200.pyscf.basis import BasisLayout
    out = {}
    for wl, dmk in (("benzene-ccpvtz", "ones"), ("taxol-svp", "ones"), ("valinomycin-tzvp", "decay"), ("valinomycin-tzvpp", "ones")):
        mol, label = build_mol(wl)
        lay = BasisLayout.from_mol(mol, alignment=4)
        eng = lay.engine()
        dm_dev = torch.as_tensor(make_dm(mol, dmk), device=dev)
        eng.q_matrix(0.0)
        ms = time_builds(eng, dm_dev, 1, 1 if mol.nao > 1500 else 3)
        counts, pw, _ = eng.last_stats()
        fl = total_flops(counts, pw)
        out["%s/dm=%s" % (wl, dmk)] = {"s_per_build": ms * 1e-3, "quartets": int(counts.sum()), "tflops": fl / (ms * 1e-3) / 1e12,
                                       "frac_of_fp64_peak": fl / (ms * 1e-3) / 1e12 / peak_probe}
        if wl == "valinomycin-tzvpp":      # config 5: FP64 vs mixed
            vj64, vk64 = eng.get_jk(dm_dev, hermi=1)
            out["%s/dm=%s" % (wl, dmk)]["mixed_precision"] = mixed_leg(eng, dm_dev, vj64, vk64)
            del vj64, vk64
        lay._cache.clear()
        del eng
        torch.cuda.empty_cache()
    return out


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "jqc-kernels"])
    ap.add_argument("--workload", default="valinomycin-tzvp", choices=sorted(WORKLOADS))
    ap.add_argument("--dm", default="ones", choices=["ones", "decay"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other configs and the reference-kernel leg")
    ap.add_argument("--class-profile", default=None, help="write the per-class device-time table to this file")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from joltqc_b200.pyscf.basis import BasisLayout
    mol, label = build_mol(args.workload)
    lay = BasisLayout.from_mol(mol, alignment=4)
    dm = make_dm(mol, args.dm)
    config = _config(args, mol, lay, label, world)

    if args.impl == "reference":
        # The reference's CPU path for get_jk is PySCF/libcint, which is not installed; the in-repo
        # CPU restatement (oracle port) is timed instead, on all host cores (set explicitly: torchrun
        # exports OMP_NUM_THREADS=1).  Rank 0 only.
        if rank == 0:
            run_reference_arm(args, lay, dm, config)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: joltqc_b200 has no CPU J/K path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    if args.impl == "jqc-kernels":
        if rank != 0:
            return
        eng = lay.engine()
        dm_dev = torch.as_tensor(dm, device=dev)
        eng.q_matrix(0.0)
        leg = ref_kernels_leg(lay, eng, dm_dev, reps=max(1, args.steps))
        if "unavailable" in leg:
            _emit({"impl": "jqc-kernels", "unavailable": leg["unavailable"]})
            return
        _emit({"impl": "jqc-kernels", "metric": "fp64_jk_build_time", "value": leg["value"], "unit": "s/iter", "n_gpus": 1,
               "steps": max(1, args.steps), "warmup": 1, "ms_per_step": leg["value"] * 1e3, "higher_is_better": False,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "ref_kernels": leg})
        return

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = lay.engine()
    if world > 1:
        eng.enable_sharding(rank, world)
    dm_dev = torch.as_tensor(dm, device=dev)
    dm_pin = torch.as_tensor(dm).pin_memory()
    out_pin = (torch.empty_like(dm_pin).pin_memory(), torch.empty_like(dm_pin).pin_memory())
    t0 = time.perf_counter()
    eng.q_matrix(0.0)
    torch.cuda.synchronize()
    t_schwarz = time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return eng.get_jk(dm_dev, hermi=1)

    def step_e2e():
        if world == 1:
            # the host-buffer C entry point: H2D of D, the build, D2H of J and K inside jqc_get_jk_host
            return eng.get_jk_host(dm_pin.numpy(), hermi=1, out=(out_pin[0].numpy(), out_pin[1].numpy()))
        d = dm_pin.to(dev, non_blocking=True)           # every rank receives D from its host
        vj, vk = eng.get_jk(d, hermi=1)
        if rank == 0:
            out_pin[0].copy_(vj, non_blocking=True)
            out_pin[1].copy_(vk, non_blocking=True)
        torch.cuda.synchronize()
        return None

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        vj, vk = step_device()
    ev1.record()
    barrier()
    ms_local = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    # build-only time of this rank (before the all_reduce): one extra step, timed around build_partial
    torch.cuda.synchronize()
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eb0.record()
    eng.build_partial(dm_dev, hermi=1)
    eb1.record()
    barrier()
    ms_build_local = eb0.elapsed_time(eb1)
    counts, pw, launches = eng.last_stats()
    flops_local = total_flops(counts, pw)
    t = torch.tensor([ms_local, flops_local, float(counts.sum()), ms_build_local], dtype=torch.float64, device=dev)
    ms = ms_local
    per_rank = None
    if dist is not None:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        ms = max(float(x[0]) for x in allt)
        flops_all = sum(float(x[1]) for x in allt)
        nq_all = sum(float(x[2]) for x in allt)
        per_rank = {"partial_build_ms": [round(float(x[3]), 2) for x in allt], "quartets": [int(float(x[2])) for x in allt],
                    "step_ms": [round(float(x[0]), 2) for x in allt]}
    else:
        flops_all, nq_all = flops_local, float(counts.sum())

    # end to end through host buffers
    step_e2e()
    barrier()
    n_e2e = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])

    # checksum of the result for the record (identical inputs -> comparable across N)
    chk = float(vj.double().sum().item()), float(vk.double().sum().item())

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # per-class device time (separate profiled pass, outside the timed region; 1 GPU share)
    rows = []
    c2 = pw2 = class_ms = None
    if world == 1:
        eng.set_profiling(True)
        step_device()
        class_ms = eng.last_class_ms()
        c2, pw2, _ = eng.last_stats()
        eng.set_profiling(False)
    for key in (np.nonzero(c2)[0] if c2 is not None else []):
        fe, fd = flops_per_class(int(key))
        fl = float(pw2[key]) * fe + float(c2[key]) * fd
        rows.append((float(class_ms[key]), int(key), int(c2[key]), fl))
    rows.sort(reverse=True)

    from joltqc_b200.backend.engine import fp64_peak_probe
    peak_probe, _ = fp64_peak_probe(local_rank)
    sm_mhz = clocks.get("sm_mhz") or 0.0
    nsm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peak_clock = nsm * 128 * sm_mhz * 1e6 / 1e12 if sm_mhz else None
    achieved = flops_all / (ms * 1e-3) / 1e12
    peak_total = peak_probe * max(world, 1)
    top = [{"class": "(%d%d|%d%d)" % (k // 125, k // 25 % 5, k // 5 % 5, k % 5), "ms": round(m, 3), "quartets": q,
            "tflops": round(fl / (m * 1e-3) / 1e12, 3) if m > 0 else None} for m, k, q, fl in rows[:8]]
    if args.class_profile:
        with open(args.class_profile, "w") as f:
            f.write("class,ms,quartets,alg_flops,tflops,frac_of_probe_peak\n")
            for m, k, q, fl in rows:
                tf = fl / (m * 1e-3) / 1e12 if m > 0 else 0.0
                f.write("(%d%d|%d%d),%.4f,%d,%.4e,%.4f,%.4f\n" % (k // 125, k // 25 % 5, k // 5 % 5, k % 5, m, q, fl, tf, tf / peak_probe))

    line = {
        "metric": "fp64_jk_build_time", "value": ms * 1e-3, "unit": "s/iter", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_total, "unit": "TFLOP/s", "frac": achieved / peak_total,
                     "traffic": None,
                     "peak_source": "DFMA probe kernel measured in this run (%.2f TFLOP/s per GPU); nominal SMs x 128 x clock = %s TFLOP/s at the sampled %.0f MHz; MEASURED_PEAKS.json has no FP64 entry" % (peak_probe, ("%.2f" % peak_clock) if peak_clock else "n/a", sm_mhz),
                     "algorithmic_flops_per_build": flops_all, "quartets_per_build": nq_all,
                     "kernel": "whole get_jk (Rys J/K kernels incl. in-kernel screening, AO transforms); per-class shares in profiles/",
                     "top_classes": top},
        "e2e": {"value": e2e_s, "unit": "s/iter", "h2d_bytes_per_step": int(dm_pin.numel() * 8) * max(world, 1),
                "d2h_bytes_per_step": int(2 * dm_pin.numel() * 8),
                "path": "jqc_get_jk_host (pinned host buffers)" if world == 1 else "per-rank H2D of D + sharded build + all_reduce + D2H on rank 0"},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "schwarz_setup_s": t_schwarz,
        "checksum": {"sum_J": chk[0], "sum_K": chk[1]},
    }
    if per_rank is not None:
        line["per_rank"] = per_rank
    # Auxiliary legs (comparisons beside the headline, all outside the timed region): a failure in one of them
    # is recorded in the line instead of losing the measurement above.
    def leg(name, fn):
        try:
            line[name] = fn()
        except Exception as exc:                       # noqa: BLE001 - reported, not swallowed
            line[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            print("bench.py: leg %r failed: %r" % (name, exc), file=sys.stderr)

    if world == 1:
        leg("mixed_precision", lambda: mixed_leg(eng, dm_dev, vj, vk))
    if world == 1 and not args.no_extras:
        leg("ref_kernels", lambda: ref_kernels_leg(lay, eng, dm_dev))
        if "value" in line["ref_kernels"]:
            line["ref_kernels"]["speedup_of_this_engine"] = line["ref_kernels"]["value"] / (ms * 1e-3)
        leg("extra", lambda: extra_configs(peak_probe, dev))
    if not args.no_cpu_baseline and world == 1:
        def cpu_leg():
            res = cpu_port_baseline(lay, dm)
            est = res["seconds_sample"] * res["stride"]
            return {"value": est, "unit": "s/iter", "cores": res["cores"], "kind": "port",
                    "sample": f"every {res['stride']}th (ij) shell pair of the same build "
                              f"({res['quartets_sample']} quartets in {res['seconds_sample']:.2f} s), extrapolated x{res['stride']}",
                    "anchor": cpu_anchor()}
        leg("cpu_baseline", cpu_leg)
    _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
