"""Summarise a single-pass ncu launch list (hardware counters only, no replay) per kernel instantiation:
device time, time-weighted FP64-pipe utilisation (sm__pipe_fp64_cycles_active, the hardware-side fraction of
the FP64 peak), issue-slot utilisation, resident warps, registers.

    ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,\
sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
        -k regex:jk_ --csv --log-file launches.csv python tools/one_build.py valinomycin-tzvp 1
    python tools/ncu_pipe_classes.py launches.csv > per_kernel.csv
"""
import csv
import re
import sys
from collections import defaultdict

UNIT_MS = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
FP64 = "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"
WARPS = "sm__warps_active.avg.pct_of_peak_sustained_active"
ISSUE = "smsp__issue_active.avg.pct_of_peak_sustained_active"
REGS = "launch__registers_per_thread"
TIME = "gpu__time_duration.sum"


def main(path):
    launch = {}
    for r in csv.reader(open(path)):
        if len(r) > 10 and r[0].isdigit():
            launch.setdefault((r[0], r[4]), {})[r[-3]] = (float(r[-1].replace(",", "")), r[-2])
    agg = defaultdict(lambda: defaultdict(float))
    for (_, name), d in launch.items():
        t, unit = d.get(TIME, (0.0, "ns"))
        ms = t * UNIT_MS.get(unit, 1e-6)
        a = agg[re.sub(r"\(.*\)$", "", name).replace("void ", "").replace("jqc::", "")]
        a["n"] += 1
        a["ms"] += ms
        for m in (FP64, WARPS, ISSUE):
            a[m] += ms * d.get(m, (0.0,))[0]
        a["regs"] = max(a["regs"], d.get(REGS, (0.0,))[0])
    tot = sum(a["ms"] for a in agg.values())
    print("kernel,launches,ms,share_of_listed_time,fp64_pipe_active_pct,issue_active_pct,warps_active_pct,registers")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        if a["ms"] <= 0:
            continue
        print('"%s",%d,%.3f,%.4f,%.1f,%.1f,%.1f,%d' % (name, a["n"], a["ms"], a["ms"] / tot, a[FP64] / a["ms"], a[ISSUE] / a["ms"],
                                                      a[WARPS] / a["ms"], int(a["regs"])))
    w = sum(a[FP64] for a in agg.values()) / tot if tot else 0.0
    print('"ALL (time-weighted)",%d,%.3f,1.0000,%.1f,%.1f,%.1f,' % (sum(a["n"] for a in agg.values()), tot, w,
                                                                   sum(a[ISSUE] for a in agg.values()) / tot if tot else 0.0,
                                                                   sum(a[WARPS] for a in agg.values()) / tot if tot else 0.0))


if __name__ == "__main__":
    main(sys.argv[1])
