#!/usr/bin/env python
"""Per-source-line profile of one kernel from an .ncu-rep captured with --import-source on:
warp instructions executed (share), threads per instruction, stall samples (share), long-scoreboard share.
usage: src_profile.py <rep> [top=45]      (runs `ncu --page source --print-source cuda,sass --csv` here, no GPU)"""
import collections
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
lines = {}
cur_file, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and cur_file:
        d = dict(zip(hdr[4:], r[4:]))   # the two 'Source' columns collide; metrics start at column 4
        try:
            ie, te, sm = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])
            lsb = int(d.get("stall_long_sb") or 0)
        except (KeyError, ValueError):
            continue
        if ie == 0 and sm == 0:
            continue
        k = (cur_file, int(r[0]))
        a = lines.setdefault(k, [0, 0, 0, 0, r[1].strip()[:70]])
        a[0] += ie; a[1] += te; a[2] += sm; a[3] += lsb
tot = sum(v[0] for v in lines.values())
tots = sum(v[2] for v in lines.values())
thr = sum(v[1] for v in lines.values())
print(f"total warp instructions {tot}  threads/inst {thr / max(tot, 1):.2f}  stall samples {tots}")
byfile = collections.Counter()
for k, v in lines.items():
    byfile[k[0]] += v[0]
for f, c in byfile.most_common(8):
    print(f"  {f:28s} {100 * c / tot:5.1f}%")
print(f"{'file:line':26s} {'inst%':>6s} {'thr/inst':>8s} {'stall%':>6s} {'longsb%':>7s}  source")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:20]:20s}:{k[1]:4d} {100 * v[0] / tot:6.2f} {v[1] / max(v[0], 1):8.1f} {100 * v[2] / max(tots, 1):6.2f} {100 * v[3] / max(tots, 1):7.2f}  {v[4]}")
