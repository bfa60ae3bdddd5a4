#!/usr/bin/env python
"""Least-squares vertex filter on a workload's bake: CG iterations and time for both regulariser forms.
usage: ls_probe.py <c1|c2|c4|c5> [weight=0.1]"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import bench  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402

w = sys.argv[1]
weight = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
scene, blockers, min_per, requested, desc = bench.make_workload(w)
rays = bench.RAYS[w]
off, maxd = scenes.default_distances(scene)
ao = None
for energy in (1, 0):
    with api.Baker(ls_energy=energy, cg_max_iterations=50000) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(min_per, requested)
        bk.sample_instances(per, min_per, download=False)
        if ao is None:
            ao = bk.compute_ao(rays, off, maxd)
        bk.set_ao(ao)
        for rep in range(2):
            t0 = time.perf_counter()
            v = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, weight)
            dt = time.perf_counter() - t0
        tm = bk.timings()
        va = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)
        print(f"{w} ls_energy {energy} w {weight}: {tm.cg_iterations} CG iterations, {dt * 1e3:.1f} ms (filter_ms {tm.filter_ms:.1f}); vertex AO mean {np.mean([x.mean() for x in v]):.4f} "
              f"min {min(x.min() for x in v):.3f} max {max(x.max() for x in v):.3f}; averaging filter mean {np.mean([x.mean() for x in va]):.4f}", flush=True)
