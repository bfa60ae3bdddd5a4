#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv` SASS dump by CUDA source line, using the
`//## File ... line N` annotations of `nvdisasm -g` for the same kernel.
usage: line_profile.py <src.csv> <kernel.sass> [top]"""
import collections
import csv
import re
import sys

src_csv, sass, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
addr2line = {}
cur = None
inl = None
for ln in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ia, iex, ith, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r in data:
    if not r[iex].isdigit():
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    agg[key][0] += int(r[iex]); agg[key][1] += int(r[ith]); agg[key][2] += int(r[isamp] or 0)
    tot += int(r[iex])
print(f"total warp instructions {tot}")
byfile = collections.Counter()
for k, v in agg.items():
    byfile[k[0]] += v[0]
for f, c in byfile.most_common():
    print(f"  {f:20s} {100 * c / tot:5.1f}%")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:18s}:{k[1]:4d}  {100 * v[0] / tot:5.2f}%  thr/inst {v[1] / max(v[0], 1):5.1f}  samples {v[2]}")
