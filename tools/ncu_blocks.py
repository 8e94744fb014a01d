#!/usr/bin/env python
"""Basic-block view of an ncu report's SASS page: consecutive instructions with the same execution
count are merged; prints executed warp instructions, share, stall samples and average active lanes.
usage: tools/ncu_blocks.py <report.ncu-rep> [launch-index] [min-share-pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = [r[1]]
        secs.append(cur)
    elif cur is not None:
        cur.append(r)
name, hdr, body = secs[which][0], secs[which][1], secs[which][2:]
ia, isrc, ii, it, ism = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed",
                                                 "Thread Instructions Executed", "# Samples"))
base = int(body[0][ia], 16)
tot = sum(float(r[ii]) for r in body)
ts = sum(float(r[ism]) for r in body)
blocks, blk = [], None
for r in body:
    off, n, t, s = int(r[ia], 16) - base, float(r[ii]), float(r[it]), float(r[ism])
    op = r[isrc].split()[0] if r[isrc].split() else ""
    if blk and blk["n"] == n:
        blk["len"] += 1; blk["inst"] += n; blk["thr"] += t; blk["smp"] += s; blk["end"] = off; blk["ops"].append(op)
    else:
        blk = dict(start=off, end=off, n=n, len=1, inst=n, thr=t, smp=s, ops=[op])
        blocks.append(blk)
print("%s: %.4g warp instructions, %d samples" % (name, tot, ts))
for b in blocks:
    if 100 * b["inst"] / tot > min_share or 100 * b["smp"] / ts > min_share:
        print("%05x-%05x len %3d exec %10.0f inst %5.1f%% smp %5.1f%% lanes %4.1f  %s" % (
            b["start"], b["end"], b["len"], b["n"], 100 * b["inst"] / tot, 100 * b["smp"] / ts,
            b["thr"] / max(b["inst"], 1), " ".join(b["ops"][:14])))
