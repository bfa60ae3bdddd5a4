#!/usr/bin/env python
"""One compute_ao over a bounded sample range of a workload — the short command wrapped by ncu.
usage: run_range.py <c1|c2|c3|c4> <num_samples or 0 = all> [trace_kernel] [repeats] [start_fraction]
(environment: RAY_ORDER = AoBakeParams::ray_order, TRI_BATCH = AoBakeParams::tri_batch)"""
import os
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402

w = sys.argv[1]
n_req = int(sys.argv[2])
tk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
frac = float(sys.argv[5]) if len(sys.argv) > 5 else 0.37
scene, blockers, min_per, requested, desc = bench.make_workload(w)
rays = bench.RAYS[w]
off, maxd = scenes.default_distances(scene)
with api.Baker(trace_kernel=tk, ray_order=int(os.environ.get("RAY_ORDER", "0")), tri_batch=int(os.environ.get("TRI_BATCH", "0"))) as bk:
    bk.set_scene(scene, blockers)
    total, per = bk.distribute_samples(min_per, requested)
    bk.sample_instances(per, min_per, download=False)
    n = total if n_req <= 0 else min(n_req, total)
    b = int((total - n) * frac)
    for i in range(reps):
        bk.compute_ao(rays, off, maxd, download=False, begin=b, end=b + n)
        t = bk.timings()
        print(f"{w} kernel {tk} samples {n} rays {t.rays_traced} trace_ms {t.trace_ms:.3f} Grays/s {t.rays_traced / t.trace_ms / 1e6:.3f}", flush=True)
    st = bk.stats()
    print(f"bvh nodes {st.num_bvh_nodes} tris {st.num_bvh_triangles} bytes {st.bvh_bytes} build_ms {bk.timings().bvh_build_ms:.2f} depth {st.reserved[0]} blas_depth {st.reserved[1]}")
