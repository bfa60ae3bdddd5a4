#!/usr/bin/env python
"""Node visits / triangle tests / instance entries per ray of the fused kernel on the first N samples of a
workload (instrumented launch) + its time.  usage: stats_probe.py <c1|c2|c3|c4> [n_samples=200000]"""
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402

w = sys.argv[1]
n_req = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
scene, blockers, min_per, requested, desc = bench.make_workload(w)
rays = bench.RAYS[w]
off, maxd = scenes.default_distances(scene)
for stats in (True, False):
    with api.Baker(trace_kernel=2, collect_stats=stats) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(min_per, requested)
        bk.sample_instances(per, min_per, download=False)
        n = min(n_req, total)
        b = (total - n) // 2
        for _ in range(2):
            bk.compute_ao(rays, off, maxd, download=False, begin=b, end=b + n)
        t, s = bk.timings(), bk.stats()
        if stats:
            print(f"{w} stats: nodes/ray {s.node_visits / s.rays:.3f} tris/ray {s.triangle_tests / s.rays:.3f} insts/ray {s.instance_entries / s.rays:.3f}", flush=True)
        else:
            print(f"{w} time: {t.trace_ms:.2f} ms {t.rays_traced / t.trace_ms / 1e6:.2f} Grays/s  hits_sum {int(bk.hit_counts()[b:b + n].sum())}", flush=True)
