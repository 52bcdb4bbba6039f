#!/usr/bin/env python
"""Per-kernel stall summary and the hottest source lines of an ncu report:  python tools/ncu_stalls.py rep.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if pat and not pat.search(name):
        continue
    print("##", name[:100])
    for k in keys:
        if k in h:
            print(f"  {k:68s} {r[h.index(k)]:>16s} {rows[1][h.index(k)]}")
    st = [(float(r[i] or 0), c) for i, c in enumerate(h) if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
    for v, c in sorted(st, reverse=True)[:7]:
        print(f"  stall {c[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.2f} cycles/issue")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", "regex:" + sys.argv[2]] if len(sys.argv) > 2 else []),
                     capture_output=True, text=True).stdout
print(src[:200])
