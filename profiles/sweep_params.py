#!/usr/bin/env python
"""Fused-kernel time vs Baker parameters.  usage: sweep_params.py <workload> key=v1,v2,... [key=...]  (one key swept at a time
around the defaults); keys: refill_below, leaf_tris, tri_batch, node_test"""
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402

w = sys.argv[1]
scene, blockers, min_per, requested, desc = bench.make_workload(w)
rays = bench.RAYS[w]
off, maxd = scenes.default_distances(scene)
q2 = bench.sqrt_rays(rays) ** 2
for arg in sys.argv[2:]:
    key, vals = arg.split("=")
    for v in vals.split(","):
        with api.Baker(trace_kernel=2, **{key: int(v)}) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(min_per, requested)
            bk.sample_instances(per, min_per, download=False)
            n = total if w != "c4" else total // 8
            b = 0 if w != "c4" else 3 * (total // 8)
            ts = []
            for i in range(3):
                bk.compute_ao(rays, off, maxd, download=False, begin=b, end=b + n)
                ts.append(bk.timings().trace_ms)
            print(f"{w} {key}={v}: {min(ts):9.2f} ms {n * q2 / min(ts) / 1e6:6.2f} Grays/s  (bvh {bk.stats().num_bvh_nodes} nodes, build {bk.timings().bvh_build_ms:.1f} ms)", flush=True)
