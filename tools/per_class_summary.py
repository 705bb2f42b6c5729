"""Join, per angular class, the algorithmic numbers (class_times_*.csv: device time, quartets, FLOP model), the ncu
hardware counters of the kernel instantiation the class runs on (ncu_pipe_per_kernel_*.csv) and the SASS listing
index (profiles/sass/r2/INDEX.csv).  usage: per_class_summary.py class_times.csv ncu_pipe_per_kernel.csv > out.csv"""
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
times = {r["class"]: r for r in csv.DictReader(open(sys.argv[1]))}
hw = {}
for r in csv.DictReader(open(sys.argv[2])):
    m = re.match(r"(jk_\w+)<(?:double, |float, )?(\d), (\d), (\d), (\d),", r["kernel"])
    if m and "float" not in r["kernel"]:
        hw["(%s%s|%s%s)" % m.groups()[1:]] = (m.group(1), r)
sass = {r["class"]: r for r in csv.DictReader(open(os.path.join(ROOT, "profiles", "sass", "r2", "INDEX.csv")))}
print("class,kernel,ms,quartets,alg_tflops,alg_frac_of_fp64_peak,ncu_launches,ncu_ms,fp64_pipe_active_pct,issue_active_pct,"
      "warps_active_pct,registers,stack_bytes,sass_listing")
for c, t in sorted(times.items(), key=lambda kv: -float(kv[1]["ms"])):
    k, h = hw.get(c, ("", {}))
    s = sass.get(c, {})
    print(",".join([c, k or s.get("kernel", ""), t["ms"], t["quartets"], t["tflops"], t["frac_of_probe_peak"],
                    h.get("launches", ""), h.get("ms", ""), h.get("fp64_pipe_active_pct", ""), h.get("issue_active_pct", ""),
                    h.get("warps_active_pct", ""), s.get("registers", h.get("registers", "")), s.get("stack_bytes", ""),
                    ("sass/r2/" + s["file"]) if s else ""]))
