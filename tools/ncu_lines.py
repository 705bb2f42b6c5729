"""Aggregate an .ncu-rep (captured with --import-source on, built with -lineinfo) per CUDA source
line: stall samples, executed instructions, L1 tag requests (global), shared wavefronts proxy.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = defaultdict(lambda: [0, 0, 0, 0, ""])
cur_file = None; hdr = None; cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("Warp Stall Sampling (All Samples)"); i_e = hdr.index("Instructions Executed")
        i_t = hdr.index("L1 Tag Requests Global"); i_l2 = hdr.index("L2 Theoretical Sectors Global")
        continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] != "":
        cur = (cur_file, int(r[0])); agg[cur][4] = r[1].strip()[:90]
        continue
    if cur is None: continue
    a = agg[cur]
    def num(x):
        try: return int(x)
        except ValueError: return 0
    a[0] += num(r[i_s]); a[1] += num(r[i_e]); a[2] += num(r[i_t]); a[3] += num(r[i_l2])
tot = [sum(a[i] for a in agg.values()) for i in range(4)]
print("total: samples %d  instr %.3e  L1 tag req %.3e  L2 sectors %.3e" % tuple(tot))
print("%-24s %7s %7s %7s %7s  %s" % ("file:line", "stall%", "inst%", "L1req%", "L2sec%", "source"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%-24s %7.2f %7.2f %7.2f %7.2f  %s" % ("%s:%d" % k, 100 * a[0] / max(tot[0], 1), 100 * a[1] / max(tot[1], 1),
                                                 100 * a[2] / max(tot[2], 1), 100 * a[3] / max(tot[3], 1), a[4]))
