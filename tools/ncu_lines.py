#!/usr/bin/env python
"""Join an ncu report's SASS source page with nvdisasm line info and aggregate executed warp
instructions + stall samples per CUDA source line.
usage: tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top_n]"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}
cur = None
infn = False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, curk = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        curk = (r[1], [])
        sections.append(curk)
    elif curk:
        curk[1].append(r)
only = os.environ.get("NCU_LAUNCH")
for si, (k, rs) in enumerate(sections):
    if only is not None and si != int(only):
        continue
    hdr = rs[0]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(rs[1][ia], 16)
    agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
    for r in rs[1:]:
        off = int(r[ia], 16) - base
        key = line_of.get(off, ("?", 0))
        a = agg[key]
        a[0] += float(r[ii] or 0)
        a[1] += float(r[isamp] or 0)
        for i, h in stall_cols:
            a[2][h] += float(r[i] or 0)
    ti = sum(a[0] for a in agg.values())
    ts = sum(a[1] for a in agg.values())
    print("kernel:", k, " warp-inst %.3g  samples %d" % (ti, ts))
    srcs = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        f, l = key
        if f not in srcs:
            p = [p for p in glob.glob("/root/repo/**/" + f, recursive=True)]
            srcs[f] = open(p[0]).read().splitlines() if p else []
        text = srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ""
        st = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %-70s %s" % (100 * a[1] / ts, 100 * a[0] / ti, f[:14], l, text,
              " ".join("%s=%.0f%%" % (h[6:], 100 * v / max(a[1], 1)) for h, v in st)))
