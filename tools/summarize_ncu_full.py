#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one row per captured launch with duration, registers, achieved occupancy,
issue utilisation, FMA / FP64 pipe utilisation, DRAM bytes, L1/L2 hit rates and the three largest stall reasons.
Usage: summarize_ncu_full.py report.ncu-rep | report.raw.csv   (a .ncu-rep needs ncu on PATH; the raw CSV is what
`ncu -i report.ncu-rep --page raw --csv` prints -- the reports themselves are too large to bring back from the GPU box)"""
import csv
import io
import re
import subprocess
import sys


def main():
    if sys.argv[1].endswith(".csv"):
        raw = open(sys.argv[1]).read()
    else:
        raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    # ncu scales a whole column to one unit (row 2 of the CSV): bring times to microseconds and sizes to megabytes
    unit_scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "Tbyte": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    for i, h in enumerate(hdr):          # newer ncu prefixes some columns with their section ("X.Y.metric"): index by bare metric name too
        m = re.search(r"([a-z0-9_]+__[A-Za-z0-9_.]+)$", h)
        if m and m.group(1) not in idx:
            idx[m.group(1)] = i
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]

    def g(d, key, scale=1.0, fmt="%.1f"):
        try:
            return fmt % (float(d[idx[key]])*scale*unit_scale.get(units[idx[key]], 1.0))
        except Exception:
            return "-"

    print("| kernel | us | regs | warps active % | issue active % | FMA pipe % | FP64 pipe % | warp instr (M) | DRAM rd MB | DRAM wr MB | L1 hit % | L2 hit % | top stalls (warps per issue) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    seen = {}
    for d in data:
        name = re.sub(r"\(.*", "", d[idx["Kernel Name"]]).replace("void ", "").replace("mpid::", "")
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:
            continue                      # first launch of each kernel only
        st = sorted(((float(d[idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:3]
        print("| `%s` | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
            name[:60], g(d, "gpu__time_duration.sum"), g(d, "launch__registers_per_thread", fmt="%.0f"),
            g(d, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            g(d, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), g(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            g(d, "smsp__inst_executed.sum", 1e-6, "%.2f"), g(d, "dram__bytes_read.sum", 1.0, "%.2f"), g(d, "dram__bytes_write.sum", 1.0, "%.2f"),
            g(d, "l1tex__t_sector_hit_rate.pct"), g(d, "lts__t_sector_hit_rate.pct"),
            ", ".join("%s %.1f" % (n, v) for v, n in st)))


if __name__ == "__main__":
    main()
