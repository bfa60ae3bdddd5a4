#!/usr/bin/env python
"""Opcode mix and static size of one phase of the fused kernel (phases as in phase_profile.py).
usage: phase_ops.py <rep> <rays> "<phase name>" """
import collections
import importlib.util
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))
rep, rays, which = sys.argv[1], float(sys.argv[2]), sys.argv[3]
src = open(os.path.join(here, "phase_profile.py")).read()
head = src[:src.index("# An instruction inlined from a helper")]
sys.argv = ["phase_profile.py", rep, str(rays)]
g = {"__file__": os.path.join(here, "phase_profile.py")}
exec(compile(head, "phase_profile_head", "exec"), g)
rows, sass, classify = g["rows"], g["sass"], g["classify"]
txt = {}
for r in rows:
    if r and r[0] == "" and len(r) > 8 and r[2].startswith("0x"):
        txt[int(r[2], 16)] = r[3].strip()
by_addr = {}
for addr, ie, te, f, ln, smp in sass:
    e = by_addr.setdefault(addr, [ie, te, smp, []])
    e[3].append((f, ln))
phase = "prologue"
sel = []
for addr in sorted(by_addr):
    ie, te, smp, levels = by_addr[addr]
    p = None
    for want in ("aob_math.cuh", "aob_bvh.cuh", "aob_kernels.cuh"):
        for f, ln in levels:
            if f == want and p is None:
                p = classify(f, ln)
    if p is not None:
        phase = p
    if phase == which:
        sel.append((addr, ie, te, txt.get(addr, "?"), levels))
tot = sum(r[1] for r in sel)
thr = sum(r[2] for r in sel)
print(f"{which}: {len(sel)} static instructions, {tot / rays:.2f} warp-inst/ray, {thr / max(tot, 1):.1f} threads/inst, {thr / rays:.1f} thread-inst/ray")
op = collections.Counter()
for a, ie, te, t, lv in sel:
    w = t.split()
    op[(w[1] if w[0].startswith("@") else w[0]).split(".")[0]] += te
for k, v in op.most_common(22):
    print(f"  {k:12s} {100 * v / max(thr, 1):5.1f}% of the phase's thread instructions  ({v / rays:6.1f} per ray)")
