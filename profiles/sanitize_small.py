#!/usr/bin/env python
"""Small bake exercising every kernel (flatten + two-level, both AO kernels, both filters, parity
hooks) — the command wrapped by compute-sanitizer (SURVEY §4 'Tools')."""
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402

for name, (scene, blockers) in {"sphere": scenes.config1_sphere(20, 20),
                                "instanced": scenes.config4_instanced(2, 12, 12, with_ground=True)}.items():
    off, maxd = scenes.default_distances(scene)
    for tk in (1, 2):
        with api.Baker(trace_kernel=tk, collect_stats=True) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(2, 3001)
            sb = bk.sample_instances(per, 2)
            ao = bk.compute_ao(16, off, maxd)
            bk.compute_ao_interleaved(1, 3, 16, off, maxd, block_samples=64)
            bk.compute_ao(16, off, maxd, download=False)
            rays = bk.dump_rays(0, 50, 16, off, maxd)
            hit = bk.trace_rays(rays.reshape(-1, 8))
            v1 = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)
            v2 = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)
            # round 2: the NCCL entry points with a one-rank communicator, sharded uploads, both PCG products
            bk.comm_init(0, 1, api.Baker.comm_unique_id())
            bk.set_scene(scene, blockers, distributed=True)
            bk.set_samples(sb, per, distributed=True)
            bk.compute_ao_distributed(16, off, maxd)
            v3 = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1, distributed=True)
            bk.comm_destroy()
        with api.Baker(trace_kernel=tk, ray_order=1, ls_matrix_free=True, ls_energy=1, tri_batch=1, no_oversized_split=True) as bk:
            bk.set_scene(scene, blockers)
            bk.set_samples(sb, per)
            bk.compute_ao(16, off, maxd, download=False)
            v4 = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)
        print(name, "kernel", tk, "samples", total, "ao mean %.4f" % ao.mean(), "hit rate %.3f" % hit.mean(), flush=True)
print("sanitize run ok")
