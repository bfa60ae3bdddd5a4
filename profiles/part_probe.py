#!/usr/bin/env python
"""Time of ONE rank's share of a workload under the interleaved partition, on one GPU: the launch a rank of an
N-GPU job runs, against 1/N of the whole pass (tail effects of small launches).
usage: part_probe.py <workload> <num_parts>"""
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402

w, parts = sys.argv[1], int(sys.argv[2])
scene, blockers, min_per, requested, desc = bench.make_workload(w)
rays = bench.RAYS[w]
off, maxd = scenes.default_distances(scene)
with api.Baker() as bk:
    bk.set_scene(scene, blockers)
    total, per = bk.distribute_samples(min_per, requested)
    bk.sample_instances(per, min_per, download=False)
    full = []
    for _ in range(2):
        bk.compute_ao(rays, off, maxd, download=False)
        full.append(bk.timings().trace_ms)
    print(f"{w} whole pass {min(full):.2f} ms; 1/{parts} = {min(full) / parts:.2f} ms", flush=True)
    for p in (0, parts // 2, parts - 1):
        ts = []
        for _ in range(2):
            bk.compute_ao_interleaved(p, parts, rays, off, maxd)
            ts.append(bk.timings().trace_ms)
        print(f"  part {p} of {parts}: {min(ts):.2f} ms ({100 * min(ts) * parts / min(full) - 100:+.1f} % over the ideal share), {bk.timings().rays_traced} rays", flush=True)
