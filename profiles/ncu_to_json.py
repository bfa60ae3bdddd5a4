#!/usr/bin/env python
"""Counters of the fused AO kernel from an `ncu --set full` capture -> profiles/<round>/ncu_<workload>.json,
the file bench.py reads for roofline.issue / roofline.traffic (run here, no GPU needed).
usage: ncu_to_json.py <rep> <workload> <rays in the captured launch> <out.json> ["capture description"]"""
import csv
import json
import subprocess
import sys

rep, workload, rays, out = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
desc = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
best = None
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    if "k_ao_persistent" in d.get("Kernel Name", ""):
        best = d
assert best is not None, "no k_ao_persistent launch in the report"
u = dict(zip(hdr, units))


def num(k):
    return float(best[k].replace(",", ""))


def scaled(k):   # ncu prints byte counts with a unit prefix
    return num(k) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[k]]


def ms(k):
    return num(k) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u[k]]


inst = num("smsp__inst_executed.sum")
res = {
    "workload": workload, "capture": desc, "kernel": best["Kernel Name"].split("(")[0], "rays": rays,
    "duration_ms": ms("gpu__time_duration.sum"), "registers": int(num("launch__registers_per_thread")),
    "ctas_per_sm_by_registers": int(num("launch__occupancy_limit_registers")),
    "warp_inst": inst, "warp_inst_per_ray": inst / rays,
    "threads_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
    "dram_bytes": scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum"),
    "l2_bytes": 32.0 * num("lts__t_sectors.sum") if "lts__t_sectors.sum" in best else None,   # 32-byte sectors
    "long_scoreboard_stall_per_issue": num("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
}
res["dram_bytes_per_ray"] = res["dram_bytes"] / rays
res["l2_bytes_per_ray"] = res["l2_bytes"] / rays if res["l2_bytes"] is not None else None
res["Grays_per_s_under_ncu"] = rays / res["duration_ms"] / 1e6
with open(out, "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))
