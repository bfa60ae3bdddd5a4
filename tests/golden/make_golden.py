"""Regenerates tests/golden/oracle_golden.json from the oracle.  The reference shipped no
golden vectors and cannot be run, so these pin the oracle against itself
across refactors; independent anchors are the analytic tests in test_oracle_cpu.py."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_binding as ob  # noqa: E402

out = {"tea": [], "lcg": [], "halton": []}
for rounds, v0, v1 in [(2, 0, 0), (2, 1, 2), (4, 0, 0), (4, 0, 1), (4, 7, 79599), (4, 999, 123456), (16, 3, 5),
                       (2, (63 << 16) | 63, 238799), (2, (1023 << 16) | 1023, 9999999)]:
    out["tea"].append([rounds, v0, v1, hex(ob.tea(rounds, v0, v1))])
for seed in [0, 1, 0xDEADBEEF]:
    out["lcg"].append([seed, [hex(x) for x in ob.lcg_stream(seed, 6)]])
for i, b in [(1, 2), (2, 2), (3, 2), (7, 2), (1000, 2), (1, 3), (2, 3), (5, 3), (26, 3), (1000, 3)]:
    out["halton"].append([i, b, hex(int(np.float32(ob.halton(i, b)).view(np.uint32)))])
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["tea"]), "tea,", len(out["lcg"]), "lcg,", len(out["halton"]), "halton vectors")

# ---- a whole small bake: samples, rays, hit counts, vertex AO (pins the oracle across refactors) ----
from optix_prime_baking_b200 import scenes  # noqa: E402
from tests.oracle_binding import Oracle  # noqa: E402

scene, blockers = scenes.config1_sphere(8, 10)
orc = Oracle(scene, blockers)
total, per = orc.distribute_samples(2, 0)
sb = orc.sample_instances(per, 2)
off, maxd = scenes.default_distances(scene)
rays = orc.generate_rays(sb, 0, 8, 16, off, maxd)
ao, hits = orc.compute_ao(sb, 16, off, maxd)
v_area = orc.filter_area(sb, ao)[0]
v_ls = orc.filter_least_squares(sb, ao, 0.1, tol=1e-12)[0]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bake_golden.npz"),
                    tri_idx=sb.infos["tri_idx"], bary=sb.infos["bary"].view(np.uint32), dA=sb.infos["dA"].view(np.uint32),
                    positions=sb.positions.view(np.uint32), normals=sb.normals.view(np.uint32), face_normals=sb.face_normals.view(np.uint32),
                    rays_first8=rays.view(np.uint32), hits=hits, ao=ao.view(np.uint32), v_area=v_area.view(np.uint32), v_ls=v_ls,
                    offset=np.float32(off), maxdist=np.float32(maxd))
print("wrote bake_golden.npz:", total, "samples,", int(hits.sum()), "hits")
