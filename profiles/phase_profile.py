#!/usr/bin/env python
"""Where the fused kernel's warp instructions go, by PHASE, from an .ncu-rep captured with --import-source on.
Every SASS instruction is attributed to the phase of its source line; lines of the shared arithmetic helpers
(ex::mul/add/..., V3 helpers, the loads) carry no phase of their own and inherit the phase of the nearest
preceding classified instruction in address order (the inlined bodies are contiguous in SASS).
usage: phase_profile.py <rep> <rays in the captured launch>"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, rays = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))

ROOT = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))


def func_ranges(path, names):
    """line ranges of the named functions / markers in a source file: {name: (first, last)}"""
    src = open(path).read().split("\n")
    out = {}
    for name, pat in names.items():
        for i, ln in enumerate(src):
            if re.search(pat, ln):
                depth, j, seen = 0, i, False
                while j < len(src):
                    depth += src[j].count("{") - src[j].count("}")
                    seen = seen or "{" in src[j]
                    if seen and depth <= 0:
                        break
                    j += 1
                out[name] = (i + 1, j + 1)
                break
    return out


csrc = ROOT + "/optix_prime_baking_b200/csrc/"
bvh = func_ranges(csrc + "aob_bvh.cuh", {"node": r"uint32_t intersect_node8_h2(_raw)?\(", "node32": r"uint32_t intersect_node8(_raw)?\(const U4", "leafslots": r"uint32_t leaf_slots_to_prims\(",
                                         "expand": r"uint32_t expand_hit_bits\(", "tri_k": r"bool test_tri_group_k\(", "tri_sel": r"bool test_tri_group_sel\(",
                                         "tri": r"bool test_tri_group\(const F4", "sphere": r"bool sphere_may_hit\(", "rcp": r"float safe_rcp\("})
mth = func_ranges(csrc + "aob_math.cuh", {"tea": r"uint32_t tea\(", "lcg": r"uint32_t lcg\(", "rnd": r"float rnd\(", "sincos": r"void sincos2pi\(",
                                          "onb": r"Onb make_onb\(", "cosdir": r"V3 cosine_dir\(", "raydir": r"V3 ao_ray_dir\(", "rayorg": r"V3 ao_ray_origin\(",
                                          "shear": r"Shear make_shear\(", "woop_k": r"bool woop_hit_k\(", "woop_sel": r"bool woop_hit_sel\("})
ksrc = open(csrc + "aob_kernels.cuh").read().split("\n")


def kline(pat):
    for i, ln in enumerate(ksrc):
        if pat in ln:
            return i + 1
    return None


K = {"kernel_begin": kline("k_ao_persistent(BvhView bvh"), "start_queued": kline("auto start_queued = [&]()"), "refill": kline("// ------------------------------ refill"),
     "lookahead": kline("const uint32_t my = (uint32_t)__popc(want_mask & lt_mask);") or kline("if (have_item && !la_valid && pass < pass_end)"),   # (stratum-major kernel: the deal-out block generates the rays)
     "after_lookahead": kline("if (!ray_active && la_count != 0u) start_queued();") or kline("if (!ray_active && la_valid) start_queued();"),
     "traverse": kline("// ------------------------------ traverse"), "tri_block": kline("const uint32_t pm = __ballot_sync(0xffffffffu, paused);") or kline("bool hit = false;"),   # (round-1 kernel: tests in place)
     "pop": kline("// Ray end and restart are written once"), "loop_end": kline("act = __ballot_sync(0xffffffffu, ray_active);   // (unchanged") or kline("const uint32_t act = __ballot_sync(0xffffffffu, ray_active);"), "kernel_end": kline("// The rays k_ao_persistent<.., H2 = true> set aside")}


def classify(f, ln):
    if f == "aob_bvh.cuh":
        for k in ("node", "node32"):
            if k in bvh and bvh[k][0] <= ln <= bvh[k][1]:
                return "node test"
        for k in ("expand", "leafslots"):
            if k in bvh and bvh[k][0] <= ln <= bvh[k][1]:
                return "leaf-mask expansion"
        for k in ("tri_k", "tri_sel", "tri"):
            if k in bvh and bvh[k][0] <= ln <= bvh[k][1]:
                return "triangle block"
        if "sphere" in bvh and bvh["sphere"][0] <= ln <= bvh["sphere"][1]:
            return "instance entry (two-level)"
        return None
    if f == "aob_math.cuh":
        for k in ("tea", "lcg", "rnd", "sincos", "cosdir", "raydir"):
            if k in mth and mth[k][0] <= ln <= mth[k][1]:
                return "ray generation"
        for k in ("shear", "woop_k", "woop_sel"):
            if k in mth and mth[k][0] <= ln <= mth[k][1]:
                return "triangle block"
        return None
    if f == "aob_kernels.cuh":
        if K["start_queued"] and K["start_queued"] <= ln < K["start_queued"] + 12:
            return "ray start (queued ray)"
        if K["refill"] <= ln < K["lookahead"]:
            return "supply / item set-up"
        if K["lookahead"] <= ln < K["after_lookahead"]:
            return "ray generation"
        if K["after_lookahead"] <= ln < K["traverse"]:
            return "loop control"
        if K["traverse"] <= ln < K["tri_block"]:
            return "node step (dispatch, push, instance entry)"
        if K["tri_block"] <= ln < K["pop"]:
            return "triangle block"
        if K["pop"] <= ln < K["loop_end"]:
            return "pop / ray end"
        if K["loop_end"] <= ln < K["kernel_end"]:
            return "loop control"
        if K["kernel_begin"] <= ln < K["refill"]:
            return "push/pop helpers, prologue"
        return None
    return None


sass = []   # (address, inst, thread_inst, file, line)
cur_file, hdr, cur_line = None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and cur_file:
        if r[0].isdigit():
            cur_line = int(r[0])
        elif r[0] == "" and len(r) > 8 and r[2].startswith("0x"):
            d = dict(zip(hdr[4:], r[4:]))
            try:
                sass.append((int(r[2], 16), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), cur_file, cur_line, int(d["# Samples"] or 0)))
            except (KeyError, ValueError):
                pass
# An instruction inlined from a helper is listed once under every level of its inline stack (ex::mul <- woop_hit_k
# <- test_tri_group <- kernel line): group the rows by address, count the instruction once, and classify it by the
# most specific level that has a phase (helper functions first, then kernel line ranges).
by_addr = {}
for addr, ie, te, f, ln, smp in sass:
    e = by_addr.setdefault(addr, [ie, te, smp, []])
    e[3].append((f, ln))
agg = collections.defaultdict(lambda: [0, 0, 0])
phase = "prologue"
for addr in sorted(by_addr):
    ie, te, smp, levels = by_addr[addr]
    p = None
    for want in ("aob_math.cuh", "aob_bvh.cuh", "aob_kernels.cuh"):
        for f, ln in levels:
            if f == want and p is None:
                p = classify(f, ln)
    if p is not None:
        phase = p
    agg[phase][0] += ie; agg[phase][1] += te; agg[phase][2] += smp
tot = sum(v[0] for v in agg.values())
thr = sum(v[1] for v in agg.values())
smp = sum(v[2] for v in agg.values())
print(f"{tot / rays:.1f} warp instructions per ray, {thr / max(tot, 1):.2f} threads per instruction, {thr / rays:.0f} thread instructions per ray")
print(f"{'phase':46s} {'inst %':>7s} {'warp-inst/ray':>13s} {'thr/inst':>8s} {'thread-inst/ray':>15s} {'stall %':>8s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:46s} {100 * v[0] / tot:7.2f} {v[0] / rays:13.2f} {v[1] / max(v[0], 1):8.1f} {v[1] / rays:15.1f} {100 * v[2] / max(smp, 1):8.2f}")
