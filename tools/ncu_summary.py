#!/usr/bin/env python
"""Summarise an `ncu --set full` report (and optionally a launch-list csv) as markdown for profiles/.
usage: tools/ncu_summary.py <report.ncu-rep> [--launches launches.csv] [--so lib.so] > profiles/xxx.md"""
import argparse
import csv
import io
import os
import subprocess
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "duration (ncu replay: cold cache, serialised)"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe busy"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe busy"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe busy"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared wavefronts % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe busy"),
])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--launches")
    ap.add_argument("--title", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# %s" % (a.title or os.path.basename(a.report)))
    print("\nSource: `%s` (ncu --set full --clock-control none --import-source on). Numbers under ncu are"
          " cold-cache, serialised replays: use the shares and counters, not the absolute time.\n" % os.path.basename(a.report))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("## %s\n\n| metric | value |\n|---|---|" % name.strip())
        for m, label in METRICS.items():
            if m in hdr:
                i = hdr.index(m)
                print("| %s (`%s`) | %s %s |" % (label, m, r[i], units[i]))
        print()
    if a.launches:
        lr = [x for x in csv.reader(open(a.launches)) if len(x) > 10]
        h = lr[0]
        ki, vi = h.index("Kernel Name"), h.index("Metric Value")
        tot = OrderedDict()
        for x in lr[1:]:
            k = x[ki].split("(")[0].strip()
            t = tot.setdefault(k, [0, 0.0])
            t[0] += 1
            t[1] += float(x[vi].replace(",", ""))
        s = sum(v[1] for v in tot.values())
        print("## launch list (`%s`)\n\n| kernel | launches | total us | share |\n|---|---|---|---|" % os.path.basename(a.launches))
        for k, (c, t) in tot.items():
            print("| %s | %d | %.1f | %.1f %% |" % (k, c, t / 1e3, 100 * t / s))


if __name__ == "__main__":
    main()
