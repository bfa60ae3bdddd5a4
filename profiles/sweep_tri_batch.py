#!/usr/bin/env python
"""Fused-kernel time vs AoBakeParams::tri_batch (lanes that must hold leaf hits before the warp runs its
triangle block).  usage: sweep_tri_batch.py <c1|c2|c3|c4> [values, default 1,2,4,6,8,12,16]"""
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402

w = sys.argv[1]
def parse(x):   # "lanes" or "lanes:iterations"
    a = x.split(":")
    return int(a[0]) | ((int(a[1]) if len(a) > 1 else 0) << 8)


vals = [parse(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,6,8,12,16").split(",")]
scene, blockers, min_per, requested, desc = bench.make_workload(w)
rays = bench.RAYS[w]
off, maxd = scenes.default_distances(scene)
ref = None
for tb in vals:
    with api.Baker(trace_kernel=2, tri_batch=tb) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(min_per, requested)
        bk.sample_instances(per, min_per, download=False)
        n = total if w != "c4" else total // 8
        b = 0 if w != "c4" else 3 * (total // 8)      # c4: an interior eighth of the lattice
        ts = []
        for i in range(3):
            bk.compute_ao(rays, off, maxd, download=False, begin=b, end=b + n)
            ts.append(bk.timings().trace_ms)
        h = int(bk.hit_counts()[b:b + n].astype("int64").sum())
        ref = h if ref is None else ref
        q2 = bench.sqrt_rays(rays) ** 2
        print(f"{w} tri_batch {tb & 255:2d} lanes / {tb >> 8:2d} iterations  {min(ts):9.2f} ms  {n * q2 / min(ts) / 1e6:6.2f} Grays/s  hits {'same' if h == ref else 'DIFFERENT'}", flush=True)
