"""GPU parity at BASELINE.json's FULL sizes (configs[0], [2], [3], [4]) against the CPU oracle run on
the same full-size scene, in the mode the GPU picks by default (fused persistent kernel, packed-fp16 node
test on flattened scenes, TLAS/BLAS when a mesh is instanced more than once).

The oracle cannot trace 10-38 G rays in a test, so each configuration checks
  * the whole result through size-independent properties (sample budget, hit counts <= q^2,
    ao == 1 - hits/q^2, shards/parts reproduce the single pass), and
  * exact per-sample hit counts against the oracle on a seeded subset of samples whose rays are
    generated with the RNG streams of their GLOBAL sample index (>= 99.99 % per-ray agreement), and
  * for config 5 the least-squares vertex AO of the full 10 M-vertex system within 1e-3.
Config 1 (15.3 M rays) is compared whole.  (configs[1] at full size: test_gpu_parity.py.)"""
import numpy as np
import pytest

from optix_prime_baking_b200 import scenes
from optix_prime_baking_b200.ctypes_types import SampleBuffers

from .oracle_binding import Oracle

pytestmark = pytest.mark.gpu

HIT_AGREEMENT = 0.9999
VERTEX_AO_TOL = 1e-3


@pytest.fixture(scope="module")
def api():
    from optix_prime_baking_b200 import api as _api
    _api.load_library()
    return _api


def subset(sb, pick):
    sub = SampleBuffers(len(pick))
    sub.positions[...] = sb.positions[pick]
    sub.normals[...] = sb.normals[pick]
    sub.face_normals[...] = sb.face_normals[pick]
    return sub


def oracle_hits_for(orc, sb, pick, rays, off, maxd):
    """Occluded-ray counts of the samples `pick` (global indices into sb) on the oracle's BVH."""
    sub = subset(sb, pick)
    all_rays = np.concatenate([orc.generate_rays_for(sub, k, int(g), rays, off, maxd) for k, g in enumerate(pick)])
    hit = orc.trace_rays(all_rays)
    return hit.reshape(len(pick), -1).sum(axis=1).astype(np.uint32), all_rays, hit


def agreement(ohits, ghits, rays_per_sample):
    diff = np.abs(ohits.astype(np.int64) - ghits.astype(np.int64)).sum()
    return 1.0 - diff / (len(ohits) * rays_per_sample)


def test_config1_full_size_whole_array(api):
    """configs[0]: 200 x 200 sphere on the ground plane, 3 samples/face, 64 rays, averaging filter — every
    sample's hit count against the oracle, the vertex map within 1e-3, and the analytic answer."""
    scene, blockers = scenes.config1_sphere()
    off, maxd = scenes.default_distances(scene)
    rays = 64
    with api.Baker(trace_kernel=2) as bk:   # the fused persistent kernel (auto would pick the simple one below 32 M rays)
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(3, 0)
        sb = bk.sample_instances(per, 3)
        ao = bk.compute_ao(rays, off, maxd)
        hits = bk.hit_counts()
        v = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)[0]
    with api.Baker(trace_kernel=1) as bk:   # and the simple kernel agrees exactly
        bk.set_scene(scene, blockers)
        bk.sample_instances(per, 3, download=False)
        bk.compute_ao(rays, off, maxd, download=False)
        assert np.array_equal(bk.hit_counts(), hits)
    assert total == 3 * scene.num_triangles == sb.n
    orc = Oracle(scene, blockers)
    ototal, oper = orc.distribute_samples(3, 0)
    osb = orc.sample_instances(oper, 3)
    assert ototal == total and np.array_equal(osb.infos["tri_idx"], sb.infos["tri_idx"])
    assert np.array_equal(osb.positions.view(np.uint32), sb.positions.view(np.uint32))
    oao, ohits = orc.compute_ao(osb, rays, off, maxd)
    assert agreement(ohits, hits, rays) >= HIT_AGREEMENT
    same = ohits == hits
    assert np.array_equal(ao[same].view(np.uint32), oao[same].view(np.uint32))   # AO is bit-exact wherever the counts agree
    ov = orc.filter_area(osb, oao)[0]
    assert np.abs(v - ov).max() <= VERTEX_AO_TOL
    # analytic: a convex body over an infinite plane sees AO = (1 + n.up)/2; with the default maxdistance
    # (10 x the extent) shallow downward rays meet the plane beyond the cut-off, so the bake is a little brighter
    want = 0.5 * (1.0 + sb.normals[:, 1])
    assert 0.0 <= float((ao - want).mean()) < 0.03


@pytest.fixture(scope="module")
def config3(api):
    """The 20M-triangle mesh with the ground plane of configs[4], 10 M samples, resident on the GPU; the
    oracle's BVH over the same scene.  (The plane lies below the terrain: configs[2] without it differs
    only in the rays that leave downward past the rim, and is covered by the subset check below too.)"""
    scene, _ = scenes.config3_bigmesh()
    blockers = scenes.ground_blockers(scene)
    off, maxd = scenes.default_distances(scene)
    bk = api.Baker(cg_tolerance=1e-6, cg_max_iterations=20000)
    bk.set_scene(scene, blockers)
    total, per = bk.distribute_samples(0, 10_000_000)
    sb = bk.sample_instances(per, 0)
    orc = Oracle(scene, blockers)
    yield {"scene": scene, "blockers": blockers, "off": off, "maxd": maxd, "bk": bk, "total": total, "per": per, "sb": sb, "orc": orc}
    bk.close()
    orc.close()


def test_config3_full_size_sampling_and_subset_parity(api, config3):
    c = config3
    bk, sb, total, rays = c["bk"], c["sb"], c["total"], 1024
    assert total == 10_000_000 == sb.n
    # sampling: the oracle's budget rule on 20 M triangles, bit for bit
    ototal, oper = c["orc"].distribute_samples(0, 10_000_000)
    osb = c["orc"].sample_instances(oper, 0)
    assert ototal == total
    assert np.array_equal(osb.infos["tri_idx"], sb.infos["tri_idx"])
    for a, b in ((osb.positions, sb.positions), (osb.normals, sb.normals), (osb.face_normals, sb.face_normals)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    del osb
    # trace: three disjoint ranges of 65536 samples (67 M rays each) + a two-part interleaved pass over one of them
    rng = np.random.default_rng(3)
    starts = [0, total // 2 - 1234, total - 65536]
    pick = []
    ghits = []
    for s in starts:
        ao = bk.compute_ao(rays, c["off"], c["maxd"], begin=s, end=s + 65536)
        h = bk.hit_counts()[s:s + 65536]
        assert h.max() <= rays and np.array_equal(ao, (1.0 - h.astype(np.float32) / np.float32(rays)).astype(np.float32))
        p = np.sort(rng.choice(65536, size=400, replace=False))
        pick.append(p + s)
        ghits.append(h[p])
    pick, ghits = np.concatenate(pick), np.concatenate(ghits)
    ohits, _, _ = oracle_hits_for(c["orc"], sb, pick, rays, c["off"], c["maxd"])
    assert agreement(ohits, ghits, rays) >= HIT_AGREEMENT


def test_config5_full_size_bake_least_squares(api, config3):
    """configs[4]: whole bake of the 20M-triangle mesh + ground plane, 10.24 G rays, least-squares filter.
    Whole-array properties, subset hit parity, and the LS vertex AO of all 10 M vertices against the
    oracle's solve of the same system fed the same AO values."""
    c = config3
    bk, sb, total, rays = c["bk"], c["sb"], c["total"], 1024
    ao = bk.compute_ao(rays, c["off"], c["maxd"])
    hits = bk.hit_counts()
    assert hits.max() <= rays and np.array_equal(ao, (1.0 - hits.astype(np.float32) / np.float32(rays)).astype(np.float32))
    # two interleaved parts reproduce the single pass bit for bit
    acc = np.zeros(total, dtype=np.float32)
    for p in range(2):
        bk.compute_ao_interleaved(p, 2, rays, c["off"], c["maxd"])
        acc += bk.download_ao()
    assert np.array_equal(acc.view(np.uint32), ao.view(np.uint32))
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(total, size=1200, replace=False))
    ohits, _, _ = oracle_hits_for(c["orc"], sb, pick, rays, c["off"], c["maxd"])
    assert agreement(ohits, hits[pick], rays) >= HIT_AGREEMENT
    bk.set_ao(ao)
    # default regulariser: the 10 M-unknown system is strongly regularised (w R ~ 1e4 x M) and takes
    # Jacobi-PCG a few thousand iterations; a CPU solve of the same system would take minutes, so the solution is
    # checked through its residual under the ORACLE's operator.  The vertex AO comes back rounded to fp32, and that
    # 3e-8 relative noise is high-frequency — exactly what the stiff regulariser amplifies (|wR| ~ 1e4 |M| ~ 1e4 |b|):
    # the rounding alone leaves ~2e-4 here, against O(1) for a wrong solution (the small scenes of
    # test_gpu_parity.py compare the solutions themselves to 1e-3)
    v = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)[0]
    iters = bk.timings().cg_iterations
    assert 0 < iters < 20000
    assert c["orc"].ls_residual(sb, ao, [v], 0.1, per_instance=c["per"], energy=0) < 1e-3
    # the scale-free option against the oracle's own solve of that system
    with api.Baker(ls_energy=1, cg_tolerance=1e-6) as b1:
        b1.set_scene(c["scene"], c["blockers"])
        b1.set_samples(sb, c["per"])
        b1.set_ao(ao)
        v1 = b1.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)[0]
        assert 0 < b1.timings().cg_iterations < 500
    ov = c["orc"].filter_least_squares(sb, ao, 0.1, tol=1e-6, per_instance=c["per"], energy=1)[0]
    assert np.abs(v1 - ov).max() <= VERTEX_AO_TOL
    va = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)[0]
    ova = c["orc"].filter_area(sb, ao, c["per"])[0]
    assert np.abs(va - ova).max() <= VERTEX_AO_TOL


def test_config4_full_size_subset_parity(api):
    """configs[3]: 1000 instances of a 49.6k-triangle mesh, 3 samples/face/instance (148.9 M samples),
    256 rays, TLAS/BLAS.  Ranges at the lattice boundary and in its interior, exact hit counts on a
    subset against the oracle's two-level trace."""
    scene, blockers = scenes.config4_instanced()
    off, maxd = scenes.default_distances(scene)
    rays = 256
    with api.Baker() as bk:
        bk.set_scene(scene, blockers)
        assert bk.stats().two_level == 1
        total, per = bk.distribute_samples(3, 0)
        assert total == 3 * scene.num_triangles
        sb = bk.sample_instances(per, 3)
        per_inst = int(per[0])
        rng = np.random.default_rng(4)
        pick, ghits = [], []
        for inst in (0, 555, 999):       # a corner, an interior and the last instance of the 10^3 lattice
            s = inst * per_inst
            ao = bk.compute_ao(rays, off, maxd, begin=s, end=s + per_inst)
            h = bk.hit_counts()[s:s + per_inst]
            assert h.max() <= rays and np.array_equal(ao, (1.0 - h.astype(np.float32) / np.float32(rays)).astype(np.float32))
            p = np.sort(rng.choice(per_inst, size=700, replace=False))
            pick.append(p + s)
            ghits.append(h[p])
    pick, ghits = np.concatenate(pick), np.concatenate(ghits)
    orc = Oracle(scene, blockers)
    assert orc.is_two_level()
    ohits, _, _ = oracle_hits_for(orc, sb, pick, rays, off, maxd)
    assert agreement(ohits, ghits, rays) >= HIT_AGREEMENT
    # the interior instance is more occluded than the corner one
    assert ghits[700:1400].mean() > ghits[:700].mean()


def test_single_rank_communicator_and_distributed_entry_points(api):
    """The NCCL entry points with a 1-rank communicator (what bench.py drives at N = 1): comm_init,
    set_scene_distributed, set_samples_distributed, compute_ao_distributed and the distributed vertex map
    equal the plain calls."""
    scene, blockers = scenes.config1_sphere(48, 48)
    off, maxd = scenes.default_distances(scene)
    with api.Baker(trace_kernel=2) as a, api.Baker(trace_kernel=2) as b:
        a.set_scene(scene, blockers)
        total, per = a.distribute_samples(3, 0)
        sb = a.sample_instances(per, 3)
        want = a.compute_ao(64, off, maxd)
        want_v = a.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)
        b.comm_init(0, 1, api.Baker.comm_unique_id())
        b.set_scene(scene, blockers, distributed=True)
        b.set_samples(sb, per, distributed=True)
        got = b.compute_ao_distributed(64, off, maxd)
        got_v = b.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1, distributed=True)
        b.comm_destroy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(got_v[0] - want_v[0]).max() < 1e-6


@pytest.mark.parametrize("energy", [0, 1])
def test_least_squares_mid_size_solution_against_the_oracle(api, energy):
    """The solutions themselves (not residuals) at a size the oracle still solves in seconds: a 320 k-triangle warped
    heightfield with a ground plane, half of it unsampled, both regulariser forms, assembled-matrix product."""
    scene, _ = scenes.config3_bigmesh(400, seed=3)
    blockers = scenes.ground_blockers(scene)
    off, maxd = scenes.default_distances(scene)
    with api.Baker(ls_energy=energy, cg_tolerance=1e-8) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(0, scene.num_triangles // 2)
        sb = bk.sample_instances(per, 0)
        ao = bk.compute_ao(64, off, maxd)
        v = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)[0]
        iters = bk.timings().cg_iterations
        assert bk.stats().reserved[3] == 1
    orc = Oracle(scene, blockers)
    ov = orc.filter_least_squares(sb, ao, 0.1, tol=1e-8, per_instance=per, energy=energy)[0]
    assert np.abs(v - ov).max() <= VERTEX_AO_TOL
    assert abs(iters - orc.ls_iterations) <= 8 + 0.02 * orc.ls_iterations      # same method, same system: same count (the GPU checks every 8)
