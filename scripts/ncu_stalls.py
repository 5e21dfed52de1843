#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report for the judge: per launch the roofline counters, then the warp-stall
sampling totals and the top stalled SASS instructions of one launch.
  python scripts/ncu_stalls.py <report.ncu-rep> <out.md> [launch index for the stall table] [top N]"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__t_sector_hit_rate.pct", "L2_hit_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_inst_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_inst_%")]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 24
    raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = raw[0], raw[1]
    cols = [(hdr.index(k), n, units[hdr.index(k)]) for k, n in KEYS if k in hdr]
    kn, gs = hdr.index("Kernel Name"), hdr.index("Grid Size")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    secs, cur, h, name = [], [], None, None
    for r in csv.reader(txt.splitlines()):
        if r and r[0] == "Kernel Name":
            if cur:
                secs.append((name, h, cur))
            cur, h, name = [], None, r[1]
        elif r and r[0] == "Address":
            h = r
        elif h:
            cur.append(r)
    secs.append((name, h, cur))
    with open(out, "w") as f:
        f.write(f"# ncu --set full --import-source on: {rep}\n\n## counters per captured launch\n\n| kernel | grid | " + " | ".join(f"{n} [{u}]" for _, n, u in cols) + " |\n")
        f.write("|---|---|" + "---|" * len(cols) + "\n")
        for r in raw[2:]:
            f.write(f"| `{r[kn].split('(')[0][:48]}` | {r[gs]} | " + " | ".join(r[i] for i, _, _ in cols) + " |\n")
        name, h, sec = secs[which]
        ix = {x: i for i, x in enumerate(h)}
        stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        tot = {s: sum(int(r[ix[s]] or 0) for r in sec) for s in stalls}
        total = sum(int(r[ix["# Samples"]] or 0) for r in sec)
        f.write(f"\n## warp-stall sampling, launch {which}: `{name[:90]}`\n\n{total} samples.\n\n| stall reason | samples | share |\n|---|---|---|\n")
        for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            if v:
                f.write(f"| {s} | {v} | {100 * v / total:.1f} % |\n")
        f.write(f"\n## top {ntop} SASS instructions by samples (same launch)\n\n| samples | executed | instruction | main stall reasons | excessive smem wavefronts |\n|---|---|---|---|---|\n")
        for r in sorted(sec, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]:
            st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:3]
            f.write(f"| {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | `{r[ix['Source']].strip()[:70]}` | " +
                    " ".join(f"{s[6:]}={v}" for v, s in st if v) + f" | {r[ix['L1 Wavefronts Shared Excessive']]} |\n")


if __name__ == "__main__":
    main()
