"""Summarise an `ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum --csv`
launch list per kernel instantiation: device time, executed FP64 flops (2 DFMA + DMUL + DADD), hardware-side FP64 TFLOP/s.
usage: ncu_fp64_classes.py launches.csv [peak_tflops]"""
import csv, sys, re
from collections import defaultdict
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 34.2
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = defaultdict(lambda: defaultdict(float))
launch = {}
for r in rows:
    key = (r[0], r[4]); m = r[-3]; v = float(r[-1].replace(",", ""))
    launch.setdefault(key, {})[m] = (v, r[-2])
agg = defaultdict(lambda: [0.0, 0.0, 0, 0.0, 0.0])
for (lid, name), d in launch.items():
    t, unit = d.get("gpu__time_duration.sum", (0, "ns"))
    t_ms = t * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(unit, 1e-6)
    fl = 2 * d.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", (0,))[0] + d.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", (0,))[0] + d.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", (0,))[0]
    a = agg[name]; a[0] += t_ms; a[1] += fl; a[2] += 1
    a[3] = max(a[3], d.get("launch__registers_per_thread", (0,))[0]); a[4] += d.get("sm__warps_active.avg.pct_of_peak_sustained_active", (0,))[0]
print("kernel,launches,ms,executed_fp64_flops,hw_tflops,frac_of_%.1f,registers,avg_warps_active_pct" % peak)
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] <= 0: continue
    tf = a[1] / (a[0] * 1e-3) / 1e12
    short = re.sub(r"\(.*\)$", "", name).replace("void ", "")
    print('"%s",%d,%.3f,%.4e,%.3f,%.3f,%d,%.1f' % (short, a[2], a[0], a[1], tf, tf / peak, int(a[3]), a[4] / a[2]))
