"""Summarise an .ncu-rep: key raw metrics + stall samples by opcode and by source line."""
import csv, subprocess, sys, io, re
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[-1]
keep = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_global_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__pcsamp_sample_buffer")
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for h, u, v in zip(hdr, units, vals):
    if any(h == k or (k.endswith("stalled") and h.startswith(k) and h.endswith("per_issue_active.ratio")) or (k.startswith("launch__occupancy") and h.startswith(k)) for k in keep):
        try:
            if h.startswith("smsp__average_warps_issue_stalled") and float(v) < 0.3: continue
        except ValueError: pass
        print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]; data = rows[hi + 1:]
ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
c, ce = Counter(), Counter()
for r in data:
    t = r[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[isamp]); ce[op] += int(r[iex])
print("stall samples by opcode (share of samples / share of executed instructions):")
for op, n in c.most_common(12):
    print(f"  {op:8s} {100*n/tot:5.1f}% / {100*ce[op]/totex:5.1f}%")
