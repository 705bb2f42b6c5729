#!/usr/bin/env python
"""Benchmark of the FP64 direct-SCF J/K build (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME] [--dm ones|decay]

A "step" is one get_jk (J and K, hermi=1, cutoff 1e-13) of the workload.  Default workload:
config 4 of BASELINE.json — valinomycin/def2-TZVP, represented by the in-tree stand-in
geometry C62H86N2O16 (166 atoms, nao 2996; the real geometry is not available offline).
D is synthetic: all ones, exactly what the reference's own J/K benchmark feeds
(benchmarks/benchmark_jk.py:99-132), or an exponentially decaying model density (--dm decay).

Printed JSON (one line, rank 0): see the measurement contract in the task statement; `value`
is seconds per J/K build with D resident in HBM (device-timed, max over ranks), `e2e` the
same through the host-buffer C-ABI call, `roofline` the algorithmic FP64 FLOP/s of the build
(SURVEY 8d model) against the FP64 FMA peak measured on the box, `cpu_baseline` the CPU
restatement (oracle) on the host cores extrapolated from a bounded sample.
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


def cpu_port_baseline(lay, dm, budget_s=20.0):
    """Time the CPU restatement on a bounded sample (every stride-th (ij) pair) and extrapolate."""
    from oracle.oracle import OracleJK, lib
    orc = OracleJK(lay)
    cores = lib().oracle_num_threads()
    T = orc.transform()
    dmi = (T @ dm @ T.T)[None]
    t0 = time.perf_counter()
    orc.q_matrix(0.0)
    t_q = time.perf_counter() - t0
    npair = lay.nbasis * (lay.nbasis + 1) // 2
    stride = max(1, npair // 400)
    while True:
        t0 = time.perf_counter()
        orc.build_raw(dmi, 1, True, True, None, 1e-13, stride=stride, phase=stride // 2)
        dt = time.perf_counter() - t0
        nq = int(orc.last_nquartets)
        if dt > budget_s / 4 or stride == 1:
            break
        stride = max(1, int(stride / min(8.0, max(2.0, budget_s / 2 / max(dt, 1e-3)))))
    return {"seconds_sample": dt, "quartets_sample": nq, "stride": stride, "cores": cores, "schwarz_seconds": t_q,
            "counts": orc.last_counts.copy()}


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


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="valinomycin-tzvp", choices=sorted(WORKLOADS))
    ap.add_argument("--dm", default="ones", choices=["ones", "decay"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--class-profile", default=None, help="write the per-class device-time table to this file")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from joltqc_b200.pyscf.basis import BasisLayout
    mol, label = build_mol(args.workload)
    lay = BasisLayout.from_mol(mol, alignment=4)
    dm = make_dm(mol, args.dm)
    config = {"workload": label, "basis": WORKLOADS[args.workload][1], "nao": mol.nao, "nao_cart_kernel": lay.nao,
              "shells": int((~lay.pad_id).sum()), "dm": args.dm, "hermi": 1, "with_j": True, "with_k": True,
              "cutoff": 1e-13, "parallelism": f"static ij-tile interleave over {max(world, 1)} GPU(s) + 1 NCCL all_reduce",
              "l2": "working set D+J+K = %.0f MB > 126 MB L2, no explicit flush" % (3 * lay.nao**2 * 8 / 1e6)}

    if args.impl == "reference":
        # The reference's CPU path for get_jk is PySCF/libcint, which is not installed; the
        # in-repo CPU restatement (oracle port) is timed instead on all host threads.
        if rank != 0:
            return
        res = None
        for _ in range(max(1, args.warmup > 0) + 0):
            pass
        ts = []
        for _ in range(max(1, args.steps)):
            res = cpu_port_baseline(lay, dm, budget_s=20.0)
            ts.append(res["seconds_sample"])
            if sum(ts) > 120:
                break
        npair = lay.nbasis * (lay.nbasis + 1) // 2
        frac = 1.0 / res["stride"]
        est = float(np.median(ts)) / frac
        line = {"impl": "reference", "metric": "fp64_jk_build_time", "value": est, "unit": "s/iter", "n_gpus": 0,
                "steps": len(ts), "warmup": args.warmup, "ms_per_step": est * 1e3, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": est, "unit": "s/iter", "cores": res["cores"], "kind": "port",
                                 "sample": f"every {res['stride']}th (ij) shell pair of the same build "
                                           f"({res['quartets_sample']} quartets in {np.median(ts):.2f} s), extrapolated x{res['stride']}"},
                "e2e": {"value": est, "unit": "s/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(line)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: joltqc_b200 has no CPU J/K path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = lay.engine()
    if world > 1:
        eng.enable_sharding(rank, world)
    dev = torch.device("cuda", local_rank)
    dm_dev = torch.as_tensor(dm, device=dev)
    dm_pin = torch.as_tensor(dm).pin_memory()
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
        d = dm_pin.to(dev, non_blocking=True)
        vj, vk = eng.get_jk(d, hermi=1)
        if rank == 0:
            out = (vj.to("cpu", non_blocking=False), vk.to("cpu", non_blocking=False))
        else:
            out = None
        return out

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
    ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    counts, pw, launches = eng.last_stats()
    flops_local = total_flops(counts, pw)
    t = torch.tensor([ms, flops_local, float(counts.sum())], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, flops_all, nq_all = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        flops_all, nq_all = flops_local, float(counts.sum())

    # end to end through host buffers
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        step_e2e()
    e1.record()
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
                     "kernel": "whole get_jk (Rys J/K kernels + task generation + AO transforms); per-class shares in profiles/",
                     "top_classes": top},
        "e2e": {"value": e2e_s, "unit": "s/iter", "h2d_bytes_per_step": int(dm_pin.numel() * 8) * max(world, 1),
                "d2h_bytes_per_step": int(2 * dm_pin.numel() * 8)},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "schwarz_setup_s": t_schwarz,
        "checksum": {"sum_J": chk[0], "sum_K": chk[1]},
    }
    if not args.no_cpu_baseline and world == 1:
        res = cpu_port_baseline(lay, dm)
        est = res["seconds_sample"] * res["stride"]
        line["cpu_baseline"] = {"value": est, "unit": "s/iter", "cores": res["cores"], "kind": "port",
                                "sample": f"every {res['stride']}th (ij) shell pair of the same build "
                                          f"({res['quartets_sample']} quartets in {res['seconds_sample']:.2f} s), extrapolated x{res['stride']}"}
    _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
