#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.txt
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep profiles/r1_top_kernels.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__cycles_active.avg",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_read.sum"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn, mu = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        ms = v / 1e6 if r[mu].startswith("n") else (v / 1e3 if r[mu].startswith("u") else v)
        a = agg.setdefault(r[kn].split("(")[0][:80], [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ASRB_WGRAD_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none: one training step, single stream (tools/profile_step.py)\n")
        f.write(f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{a[1]:10.3f} ms {100 * a[1] / tot:5.1f}%  x{a[0]:4d}  {k}\n")
    print(open(dst).read())


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none (per launch)\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')][:110]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"{m:70s} {r[i]:>16s} {units[i]}\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
