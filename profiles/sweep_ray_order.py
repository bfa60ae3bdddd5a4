#!/usr/bin/env python
"""Fused-kernel time vs AoBakeParams::ray_order (1 = sample-major: a lane owns a sample; 2 = stratum-major: the warp deals
its item's rays out stratum by stratum) crossed with tri_batch and refill_below, plus node visits / triangle tests per ray
from one instrumented launch per order.
usage: sweep_ray_order.py <c1|c2|c3|c4|c5> [settings "order:tri_lanes:tri_wait:refill,..."]"""
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402

w = sys.argv[1]
default = "1:0:0:0,2:0:0:0,2:16:6:0,2:24:8:0,2:8:6:24,2:16:8:20"
settings = [tuple(int(v) for v in x.split(":")) for x in (sys.argv[2] if len(sys.argv) > 2 else default).split(",")]
scene, blockers, min_per, requested, desc = bench.make_workload(w)
rays = bench.RAYS[w]
q2 = bench.sqrt_rays(rays) ** 2
off, maxd = scenes.default_distances(scene)
ref = None
for order, lanes, wait, refill in settings:
    tb = lanes | (wait << 8)
    with api.Baker(trace_kernel=2, ray_order=order, tri_batch=tb, refill_below=refill) as bk:
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
        print(f"{w} ray_order {order} tri_batch {lanes:2d}/{wait:2d} refill_below {refill:2d}  {min(ts):9.2f} ms  "
              f"{n * q2 / min(ts) / 1e6:6.2f} Grays/s  hits {'same' if h == ref else 'DIFFERENT'}", flush=True)
for order in sorted({s[0] for s in settings}):
    with api.Baker(trace_kernel=2, ray_order=order, collect_stats=True) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(min_per, requested)
        bk.sample_instances(per, min_per, download=False)
        n = min(total, 1 << 20)
        b = 0 if w != "c4" else 3 * (total // 8)
        bk.compute_ao(rays, off, maxd, download=False, begin=b, end=b + n)
        st = bk.stats()
        print(f"{w} ray_order {order}: {st.node_visits / st.rays:.2f} node visits, {st.triangle_tests / st.rays:.3f} triangle tests, "
              f"{st.instance_entries / st.rays:.3f} instance entries per ray ({n} samples)", flush=True)
