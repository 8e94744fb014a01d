#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of tools/run_modes.py: DRAM bytes, issue-slot use and
instruction count of the update launch (what bench.py's roofline object quotes as traffic / frac_dram / bound).
usage: tools/ncu_traffic.py <report.ncu-rep> "<workload description>" > profiles/traffic.json"""
import csv
import io
import json
import subprocess
import sys

rep, workload = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
out = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if "dbot_particle_kernel<(int)1" not in name and "dbot_particle_kernel<1" not in name:
        continue
    f = lambda m: float(r[hdr.index(m)].replace(",", ""))  # noqa: E731
    unit = lambda m: rows[1][hdr.index(m)]  # noqa: E731
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = f("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")]
    wr = f("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
    out = {"kernel": "dbot_particle_kernel<UPDATE>", "workload": workload,
           "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
           "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "dram_pct_of_peak_ncu": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "warp_instructions_per_launch": f("smsp__inst_executed.sum"),
           "source": "ncu --set full --clock-control none (%s), dram__bytes_read.sum + dram__bytes_write.sum of one "
                     "update launch" % rep.split("/")[-1]}
    break
print(json.dumps(out, indent=1))
