#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small tracked summaries under profiles/.
  launches  <launches.csv> <out.md>     : per-kernel totals + share of the step from a --metrics gpu__time_duration.sum pass
  full      <report.ncu-rep> <out.md>   : per-launch table of the counters the roofline uses from a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"), ("lts__t_sector_hit_rate.pct", "L2_hit_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2_%peak"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem")]


def launches(src, out):
    rows = list(csv.DictReader([l for l in open(src) if l.startswith('"')]))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({src}; gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.2f} % |\n")
        f.write(f"\n{len(rows)} launches, {tot / 1e3:.1f} us total.\n\nLast launches (one evaluation step):\n\n| kernel | grid | ns |\n|---|---|---|\n")
        for r in rows[-24:]:
            f.write(f"| `{r['Kernel Name'].split('(')[0]}` | {r['Grid Size']} | {r['Metric Value']} |\n")


def full(src, out):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n, units[hdr.index(k)]) for k, n in KEYS if k in hdr]
    kn, gs = hdr.index("Kernel Name"), hdr.index("Grid Size")
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n| kernel | grid | " + " | ".join(f"{n} [{u}]" for _, n, u in cols) + " |\n")
        f.write("|---|---|" + "---|" * len(cols) + "\n")
        for r in rows[2:]:
            f.write(f"| `{r[kn].split('(')[0]}` | {r[gs]} | " + " | ".join(r[i] for i, _, _ in cols) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
