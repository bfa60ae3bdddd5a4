"""GPU parity tests: the CUDA path (through the C-ABI, ctypes) against the CPU oracle on the
same seeded inputs.  Tolerances are BASELINE.json's: bit-exact sample indices, >= 99.99 %
per-ray hit/miss agreement on identical ray sets with disagreements only next to an edge or
a t bound, per-vertex AO within 1e-3."""
import numpy as np
import pytest

from optix_prime_baking_b200 import scenes
from optix_prime_baking_b200.scenes import Instance, Mesh, Scene

from .oracle_binding import Oracle

pytestmark = pytest.mark.gpu

HIT_AGREEMENT = 0.9999      # north_star: >= 99.99 % per-ray agreement
EDGE_MARGIN = 1e-6 * 50     # disagreements must sit within ~1e-6 (relative) of an edge / t bound;
                            # margin is measured in barycentric units of the sheared triangle, allow slack
VERTEX_AO_TOL = 1e-3


@pytest.fixture(scope="module")
def api():
    from optix_prime_baking_b200 import api as _api
    _api.load_library()
    return _api


def small_scenes():
    out = {}
    out["sphere_ground"] = scenes.config1_sphere(48, 48)
    out["heightfield"] = scenes.config2_heightfield(96, seed=1)
    sc, _ = scenes.config3_bigmesh(80, seed=3)
    out["warped_ground"] = (sc, scenes.ground_blockers(sc))
    out["instanced"] = scenes.config4_instanced(grid=2, stacks=20, slices=20, with_ground=True)
    # one instance under a non-uniform scale + shear + rotation + translation (normals go through the
    # inverse transpose; flattened to world space by the AUTO rule)
    xf = np.array([[1.7, 0.3, 0.0, 2.0], [-0.2, 0.6, 0.1, -1.0], [0.4, 0.0, 1.2, 0.5], [0, 0, 0, 1]], dtype=np.float32)
    sc = Scene([scenes.uv_sphere(24, 24, displace=0.1, seed=9)], [Instance(0, xf, 42)])
    out["sheared"] = (sc, scenes.ground_blockers(sc))
    # two huge triangles carrying tens of thousands of samples each
    sc = Scene([scenes.ground_plane([-1, 0, -1], [1, 0, 1], 1, 3.0, 0.0)], [Instance(0)])
    out["two_big_triangles"] = (sc, Scene([], []))
    return out


SCENES = small_scenes()


def assert_samples_equal(a, b):
    assert a.n == b.n
    assert np.array_equal(a.infos["tri_idx"], b.infos["tri_idx"]), "tri_idx must be bit-exact"
    for name in ("bary", "dA"):
        assert np.array_equal(a.infos[name].view(np.uint32), b.infos[name].view(np.uint32)), name
    for x, y, name in ((a.positions, b.positions, "positions"), (a.normals, b.normals, "normals"),
                       (a.face_normals, b.face_normals, "face_normals")):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), name


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("min_per_tri,requested", [(3, 0), (0, 5000), (1, 20011)])
def test_sampling_bit_exact(api, name, min_per_tri, requested):
    scene, blockers = SCENES[name]
    orc = Oracle(scene, blockers)
    ototal, oper = orc.distribute_samples(min_per_tri, requested)
    osb = orc.sample_instances(oper, min_per_tri)
    with api.Baker() as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(min_per_tri, requested)
        assert total == ototal
        assert np.array_equal(per, oper)
        sb = bk.sample_instances(per, min_per_tri)
    assert_samples_equal(sb, osb)


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("rays", [16, 36, 64, 100])   # 36 and 100: q not a power of two (IEEE-division path)
def test_rays_bit_exact(api, name, rays):
    scene, blockers = SCENES[name]
    off, maxd = scenes.default_distances(scene)
    orc = Oracle(scene, blockers)
    _, per = orc.distribute_samples(1, 0)
    osb = orc.sample_instances(per, 1)
    n = min(osb.n, 4000)
    with api.Baker() as bk:
        bk.set_scene(scene, blockers)
        bk.set_samples(osb, per)
        got = bk.dump_rays(0, n, rays, off, maxd)
    want = orc.generate_rays(osb, 0, n, rays, off, maxd)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def check_hits(orc, rays, got, want):
    mism = np.nonzero(got != want)[0]
    agree = 1.0 - len(mism) / max(len(rays), 1)
    assert agree >= HIT_AGREEMENT, f"agreement {agree}"
    if len(mism):
        m = orc.ray_margin(rays[mism])
        assert (m <= EDGE_MARGIN).all(), f"disagreeing rays are not edge cases: margins {m[:10]}"
    return agree


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_trace_rays_parity(api, name, mode):
    scene, blockers = SCENES[name]
    off, maxd = scenes.default_distances(scene)
    orc = Oracle(scene, blockers, mode)
    _, per = orc.distribute_samples(1, 0)
    osb = orc.sample_instances(per, 1)
    rng = np.random.default_rng(7)
    pick = np.sort(rng.choice(osb.n, size=min(osb.n, 2500), replace=False))
    rays = np.concatenate([orc.generate_rays(osb, int(g), int(g) + 1, 16, off, maxd).reshape(-1, 8) for g in pick])
    lo, hi = scene.world_bbox()
    ext = float((hi - lo).max())
    rnd = np.zeros((30000, 8), dtype=np.float32)
    rnd[:, 0:3] = rng.uniform(lo - 0.2 * ext, hi + 0.2 * ext, (30000, 3))
    d = rng.normal(size=(30000, 3))
    rnd[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rnd[:, 3] = rng.uniform(0, 0.01 * ext, 30000)
    rnd[:, 7] = rng.uniform(0.05 * ext, 3 * ext, 30000)
    # axis-aligned and degenerate directions exercise the zero-component paths
    ax = np.zeros((600, 8), dtype=np.float32)
    ax[:, 0:3] = rng.uniform(lo, hi, (600, 3))
    ax[np.arange(600), 4 + (np.arange(600) % 3)] = np.where(np.arange(600) % 2, 1.0, -1.0)
    ax[:, 7] = 10 * ext
    rays = np.ascontiguousarray(np.concatenate([rays, rnd, ax]), dtype=np.float32)
    want = orc.trace_rays(rays)
    with api.Baker(instancing_mode=mode) as bk:
        bk.set_scene(scene, blockers)
        assert bool(bk.stats().two_level) == orc.is_two_level()
        got = bk.trace_rays(rays)
    if name != "two_big_triangles":
        assert 0.02 < want.mean() < 0.98
    check_hits(orc, rays, got, want)


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("trace_kernel", [1, 2])
def test_compute_ao_and_vertex_maps(api, name, trace_kernel):
    scene, blockers = SCENES[name]
    off, maxd = scenes.default_distances(scene)
    rays = 64
    orc = Oracle(scene, blockers)
    with api.Baker(trace_kernel=trace_kernel, collect_stats=True) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(2, 0)
        sb = bk.sample_instances(per, 2)
        ao = bk.compute_ao(rays, off, maxd)
        hits = bk.hit_counts()
        st = bk.stats()
        assert st.rays == total * rays and st.node_visits > 0
        v_area = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)
        v_ls = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)
        # sharded ranges reproduce the full result bit for bit (multi-GPU contract)
        cut = total // 3
        a0 = bk.compute_ao(rays, off, maxd, begin=0, end=cut)
        a1 = bk.compute_ao(rays, off, maxd, begin=cut, end=total)
        assert np.array_equal(np.concatenate([a0, a1]).view(np.uint32), ao.view(np.uint32))
    oao, ohits = orc.compute_ao(sb, rays, off, maxd)
    diff_rays = np.abs(hits.astype(np.int64) - ohits.astype(np.int64)).sum()
    assert 1.0 - diff_rays / (total * rays) >= HIT_AGREEMENT
    assert np.abs(ao - oao).max() <= 2.0 / rays
    same = hits == ohits
    assert np.array_equal(ao[same].view(np.uint32), oao[same].view(np.uint32))
    o_area = orc.filter_area(sb, ao, per)
    o_ls = orc.filter_least_squares(sb, ao, 0.1, per_instance=per)
    for i in range(len(scene.instances)):
        assert np.abs(v_area[i] - o_area[i]).max() <= VERTEX_AO_TOL
        assert np.abs(v_ls[i] - o_ls[i]).max() <= VERTEX_AO_TOL


@pytest.mark.parametrize("name", ["sphere_ground", "heightfield"])
def test_fp16_and_fp32_node_tests_give_identical_hit_counts(api, name):
    """Both box tests are conservative and the triangle test decides, so the per-sample hit counts of
    the fused kernel must not depend on which one ran — including the rays the fp16 test hands to the
    deferred fp32 launch (axis-parallel directions are forced here by axis-aligned normals)."""
    scene, blockers = SCENES[name]
    off, maxd = scenes.default_distances(scene)
    rays = 256
    res = []
    for node_test in (0, 1):
        with api.Baker(trace_kernel=2, node_test=node_test) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(1, 0)
            bk.sample_instances(per, 1, download=False)
            bk.compute_ao(rays, off, maxd, download=False)
            res.append((bk.hit_counts(), bk.stats().reserved[2]))
    assert np.array_equal(res[0][0], res[1][0])
    assert res[0][1] > 0 and res[1][1] == 0          # the fp16 build did defer some rays, the fp32 build none
    # a deferred-ray list that is too small must not lose rays: the launch is repeated in fp32
    with api.Baker(trace_kernel=2, deferred_capacity=4) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(1, 0)
        bk.sample_instances(per, 1, download=False)
        bk.compute_ao(rays, off, maxd, download=False)
        assert np.array_equal(bk.hit_counts(), res[0][0])


def test_scaled_instances_under_a_tlas(api):
    """Non-rigid instance transforms: object-space rays are not unit length (the one-ray-per-thread
    kernels then take the fp32 box test per ray; the fused two-level kernel is fp32 throughout)."""
    base, _ = scenes.config4_instanced(grid=2, stacks=12, slices=12, with_ground=False)
    insts = []
    for k, inst in enumerate(base.instances):
        xf = inst.xform.copy()
        xf[:3, :3] = xf[:3, :3] @ np.diag([1.0 + 0.3 * k, 1.0, 0.5]).astype(np.float32)    # non-uniform scale
        insts.append(Instance(inst.mesh_index, xf))
    scene = Scene(base.meshes, insts)
    off, maxd = scenes.default_distances(scene)
    orc = Oracle(scene, None, 2)
    for trace_kernel in (1, 2):
        with api.Baker(trace_kernel=trace_kernel, instancing_mode=api.INSTANCING_TWO_LEVEL) as bk:
            bk.set_scene(scene)
            total, per = bk.distribute_samples(2, 0)
            sb = bk.sample_instances(per, 2)
            bk.compute_ao(64, off, maxd, download=False)
            hits = bk.hit_counts()
        _, ohits = orc.compute_ao(sb, 64, off, maxd)
        assert 1.0 - np.abs(hits.astype(np.int64) - ohits.astype(np.int64)).sum() / (total * 64) >= HIT_AGREEMENT


def test_whole_bake_golden(api):
    """The CUDA path against the committed fixture tests/golden/bake_golden.npz: bit-exact samples,
    rays and (up to one edge-case ray) hit counts; vertex AO of the golden AO values within 1e-6 / 1e-4."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bake_golden.npz"))
    scene, blockers = scenes.config1_sphere(8, 10)
    off, maxd = scenes.default_distances(scene)
    for trace_kernel in (1, 2):
        with api.Baker(trace_kernel=trace_kernel) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(2, 0)
            sb = bk.sample_instances(per, 2)
            assert np.array_equal(sb.infos["tri_idx"], g["tri_idx"])
            for name, arr in [("bary", sb.infos["bary"]), ("dA", sb.infos["dA"]), ("positions", sb.positions), ("normals", sb.normals),
                              ("face_normals", sb.face_normals)]:
                assert np.array_equal(np.ascontiguousarray(arr).view(np.uint32), g[name]), name
            assert np.array_equal(bk.dump_rays(0, 8, 16, off, maxd).view(np.uint32).reshape(g["rays_first8"].shape), g["rays_first8"])
            ao = bk.compute_ao(16, off, maxd)
            hits = bk.hit_counts()
            same = hits == g["hits"]
            assert np.abs(hits.astype(np.int64) - g["hits"].astype(np.int64)).sum() <= 1      # 4480 rays: at most one edge case
            assert np.array_equal(ao[same].view(np.uint32), g["ao"][same])
            bk.set_ao(g["ao"].view(np.float32))                                                # vertex maps of the golden AO
            assert np.abs(bk.map_ao_to_vertices(api.FILTER_AREA_BASED)[0] - g["v_area"].view(np.float32)).max() < 1e-6
            assert np.abs(bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)[0] - g["v_ls"]).max() < 1e-4      # CG stops at 1e-6 relative residual


def test_analytic_sphere_over_plane(api):
    """Known answer, no oracle: a convex body over an (effectively) infinite plane has
    AO(n) = (1 + n.up) / 2."""
    scene, _ = scenes.config1_sphere(40, 40)
    blockers = scenes.ground_blockers(scene, 1, 1.0e6, 0.03)
    with api.Baker() as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(3, 0)
        sb = bk.sample_instances(per, 3)
        ao = bk.compute_ao(1024, 0.02, 1.0e7)
    expect = (1.0 + sb.normals[:, 1]) / 2.0
    assert np.abs(ao - expect).mean() < 5e-3
    assert np.abs(ao - expect).max() < 6e-2


def test_closed_box_is_fully_occluded_and_lone_triangle_is_open(api):
    # inward-facing unit cube: every ray from an inside face hits another face
    v = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], dtype=np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = []
    for a, b, c, d in quads:
        tris += [[a, b, c], [a, c, d]]
    tris = np.array(tris, dtype=np.uint32)
    center = np.array([0.5, 0.5, 0.5], dtype=np.float32)
    for t in tris:  # orient inward
        n = np.cross(v[t[1]] - v[t[0]], v[t[2]] - v[t[0]])
        if np.dot(n, center - v[t[0]]) < 0:
            t[1], t[2] = t[2], t[1]
    box = Scene([Mesh(v, tris)], [Instance(0)])
    with api.Baker() as bk:
        bk.set_scene(box)
        total, per = bk.distribute_samples(8, 0)
        bk.sample_instances(per, 8, download=False)
        ao = bk.compute_ao(64, 1e-3, 100.0)
    assert np.all(ao == 0.0)
    tri = Scene([Mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float32), np.array([[0, 2, 1]], dtype=np.uint32))],
                [Instance(0)])
    with api.Baker() as bk:
        bk.set_scene(tri)
        total, per = bk.distribute_samples(16, 0)
        bk.sample_instances(per, 16, download=False)
        ao = bk.compute_ao(64, 1e-3, 100.0)
        v_area = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)
    assert np.all(ao == 1.0) and np.allclose(v_area[0], 1.0)


def test_edge_cases(api):
    scene, blockers = SCENES["sphere_ground"]
    for bad in (dict(ray_order=3), dict(ray_order=-1), dict(trace_kernel=5)):
        with pytest.raises(api.AoBakeError):       # kernel selectors outside their range are rejected at create
            api.Baker(**bad)
    with api.Baker() as bk:
        with pytest.raises(api.AoBakeError):
            bk.compute_ao(64, 0.1, 1.0)           # no scene yet
        bk.set_scene(scene, Scene([], []))        # empty blockers
        total, per = bk.distribute_samples(0, 0)  # zero samples
        assert total == 0
        sb = bk.sample_instances(per, 0)
        assert sb.n == 0
        assert bk.compute_ao(64, 0.1, 1.0).size == 0
        with pytest.raises(api.AoBakeError):
            bk.sample_instances([1], 3)           # below the per-face minimum
        with pytest.raises(api.AoBakeError):
            bk.compute_ao(0, 0.1, 1.0)
    # an empty scene with only a blocker, and a degenerate (zero-area) triangle
    v = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0]], dtype=np.float32)
    deg = Scene([Mesh(v, np.array([[0, 1, 2], [0, 1, 3]], dtype=np.uint32))], [Instance(0)])
    orc = Oracle(deg)
    with api.Baker() as bk:
        bk.set_scene(deg)
        total, per = bk.distribute_samples(2, 7)
        ototal, oper = orc.distribute_samples(2, 7)
        assert total == ototal == 7 and np.array_equal(per, oper)
        sb = bk.sample_instances(per, 2)
        osb = orc.sample_instances(oper, 2)
        assert np.array_equal(sb.infos["tri_idx"], osb.infos["tri_idx"])
        ao = bk.compute_ao(16, 1e-3, 10.0)
        assert np.isfinite(ao).all()


def test_malformed_scenes_are_errors_not_faults(api):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    good = Mesh(v, np.array([[0, 1, 2]], dtype=np.uint32))
    bad_index = Mesh(v, np.array([[0, 1, 3]], dtype=np.uint32))           # index 3 of 3 vertices
    singular = np.eye(4, dtype=np.float32)
    singular[2, 2] = 0.0
    nan_xf = np.eye(4, dtype=np.float32)
    nan_xf[0, 3] = np.nan
    with api.Baker() as bk:
        for scene, blockers in [(Scene([bad_index], [Instance(0)]), None),
                                (Scene([good], [Instance(0)]), Scene([bad_index], [Instance(0)])),
                                (Scene([good], [Instance(0, singular)]), None),
                                (Scene([good], [Instance(0, nan_xf)]), None)]:
            with pytest.raises(api.AoBakeError) as e:
                bk.set_scene(scene, blockers)
            assert e.value.code == 1   # AOBAKE_ERR_INVALID_ARGUMENT
            with pytest.raises(api.AoBakeError):
                bk.compute_ao(16, 1e-3, 1.0)      # the failed set_scene left no scene behind
        # the context is still usable afterwards (no sticky CUDA error)
        bk.set_scene(Scene([good], [Instance(0)]))
        total, per = bk.distribute_samples(4, 0)
        bk.sample_instances(per, 4)
        assert np.all(bk.compute_ao(16, 1e-3, 10.0) == 1.0)
        # an empty sample set maps to all-zero vertex AO instead of a state error
        total, per = bk.distribute_samples(0, 0)
        bk.sample_instances(per, 0)
        assert bk.compute_ao(16, 1e-3, 10.0).size == 0
        assert np.all(bk.map_ao_to_vertices(api.FILTER_AREA_BASED)[0] == 0.0)
        assert np.all(bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)[0] == 0.0)


def test_strided_vertices(api):
    scene, blockers = SCENES["sphere_ground"]
    m = scene.meshes[0]
    wide = np.zeros((len(m.vertices), 8), dtype=np.float32)
    wide[:, 0:3] = m.vertices
    wide[:, 4:7] = m.normals
    strided = Scene([Mesh.__new__(Mesh)], [Instance(0)])
    sm = strided.meshes[0]
    sm.vertices, sm.normals, sm.tris = wide[:, 0:3], wide[:, 4:7], m.tris  # views with a 32-byte stride
    with api.Baker() as a, api.Baker() as b:
        a.set_scene(scene, blockers)
        b.set_scene(strided, blockers)
        ta, pa = a.distribute_samples(1, 0)
        tb, pb = b.distribute_samples(1, 0)
        assert_samples_equal(a.sample_instances(pa, 1), b.sample_instances(pb, 1))


def test_full_size_config2_properties_and_subset_parity(api):
    """BASELINE.json configs[1] at full size (1M triangles, 3.0M samples, 256 rays/sample):
    size-independent properties on the whole result, plus exact hit-count parity with the oracle
    on a seeded subset of samples (the oracle traces the same full-size scene)."""
    scene, blockers = scenes.config2_heightfield()
    off, maxd = scenes.default_distances(scene)
    rays = 256
    with api.Baker() as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(3, 0)
        assert total == 3 * scene.num_triangles
        sb = bk.sample_instances(per, 3)
        ao = bk.compute_ao(rays, off, maxd)
        hits = bk.hit_counts()
        # shards reproduce the single pass (checksum of checksums over 4 ragged shards)
        cuts = [0, total // 5, total // 2, total - 7, total]
        parts = [bk.compute_ao(rays, off, maxd, begin=cuts[i], end=cuts[i + 1]) for i in range(4)]
        v_area = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)[0]
    assert np.array_equal(np.concatenate(parts).view(np.uint32), ao.view(np.uint32))
    assert hits.max() <= rays and np.array_equal(ao, (1.0 - hits.astype(np.float32) / np.float32(rays)).astype(np.float32))
    assert np.array_equal(np.bincount(sb.infos["tri_idx"], minlength=scene.num_triangles), np.full(scene.num_triangles, 3))
    assert 0.0 <= v_area.min() and v_area.max() <= 1.0 and abs(v_area.mean() - ao.mean()) < 0.02
    # subset parity against the oracle on the same full-size scene
    rng = np.random.default_rng(11)
    pick = np.sort(rng.choice(total, size=3000, replace=False))
    from optix_prime_baking_b200.ctypes_types import SampleBuffers
    sub = SampleBuffers(len(pick))
    sub.positions[...] = sb.positions[pick]
    sub.normals[...] = sb.normals[pick]
    sub.face_normals[...] = sb.face_normals[pick]
    orc = Oracle(scene, blockers)
    # the ray RNG is keyed by the global sample index: trace the subset one index at a time
    ohits = np.zeros(len(pick), dtype=np.uint32)
    for k, g in enumerate(pick):
        r = orc.generate_rays_for(sub, k, int(g), rays, off, maxd)
        ohits[k] = orc.trace_rays(r).sum()
    diff = np.abs(ohits.astype(np.int64) - hits[pick].astype(np.int64)).sum()
    assert 1.0 - diff / (len(pick) * rays) >= HIT_AGREEMENT


def test_least_squares_with_unsampled_regions(api):
    """< 1 sample per triangle (config 3/5 regime): the leftover rule leaves whole regions
    unsampled; zero-lumped-mass vertices are anchored (decision #7) and CG converges quickly."""
    scene, blockers = SCENES["warped_ground"]
    off, maxd = scenes.default_distances(scene)
    orc = Oracle(scene, blockers)
    with api.Baker() as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(0, scene.num_triangles // 3)
        sb = bk.sample_instances(per, 0)
        ao = bk.compute_ao(64, off, maxd)
        v_ls = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)[0]
        v_area = bk.map_ao_to_vertices(api.FILTER_AREA_BASED)[0]
        iters = bk.timings().cg_iterations
    o_ls = orc.filter_least_squares(sb, ao, 0.1, per_instance=per)[0]
    o_area = orc.filter_area(sb, ao, per)[0]
    assert np.abs(v_ls - o_ls).max() <= VERTEX_AO_TOL and np.abs(v_area - o_area).max() <= VERTEX_AO_TOL
    assert (v_area == 0).mean() > 0.3          # a large unsampled region exists
    assert iters < 500


@pytest.mark.parametrize("name", ["sphere_ground", "instanced"])
@pytest.mark.parametrize("parts,block", [(2, 64), (3, 2048), (8, 32)])
def test_interleaved_parts_sum_to_full_bake(api, name, parts, block):
    """Multi-GPU interleaved partition, exercised on one GPU: the parts' AO arrays (zeros outside a
    part's super-blocks) sum exactly to the single-launch result."""
    scene, blockers = SCENES[name]
    off, maxd = scenes.default_distances(scene)
    with api.Baker() as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(1, 10007)
        bk.sample_instances(per, 1, download=False)
        want = bk.compute_ao(64, off, maxd)
        acc = np.zeros(total, dtype=np.float32)
        owned = np.zeros(total, dtype=np.int32)
        rays = 0
        for p in range(parts):
            bk.compute_ao_interleaved(p, parts, 64, off, maxd, block_samples=block)
            part = bk.download_ao()
            rays += bk.timings().rays_traced
            idx = (np.arange(total) // block) % parts == p
            assert np.all(part[~idx] == 0.0)
            owned += idx
            acc += part
        assert rays == total * 64 and np.all(owned == 1)
        assert np.array_equal(acc.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("name", ["sphere_ground", "heightfield", "instanced"])
@pytest.mark.parametrize("rays", [1, 9, 64, 100, 1024])
def test_ray_orders_give_identical_hit_counts(api, name, rays):
    """The fused kernel traces the rays of a work item sample-major (a lane owns a sample) or stratum-major (the warp
    deals the item's rays out stratum by stratum; hits counted per sample in shared memory).  The order changes which
    rays share a warp, never a ray: per-sample hit counts must be identical, and equal to the one-ray-per-thread kernel's
    — for a sample count that is not a multiple of 32, single-ray items (rays = 1: every refill crosses an item
    boundary), non-power-of-two strata, and as interleaved parts."""
    scene, blockers = SCENES[name]
    off, maxd = scenes.default_distances(scene)
    res = {}
    for key, kw in (("simple", dict(trace_kernel=1)), ("sample_major", dict(trace_kernel=2, ray_order=1)),
                    ("stratum_major", dict(trace_kernel=2, ray_order=2))):
        with api.Baker(**kw) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(1, 10007)
            bk.sample_instances(per, 1, download=False)
            bk.compute_ao(rays, off, maxd, download=False)
            res[key] = bk.hit_counts()
            if key != "simple":
                acc = np.zeros(total, dtype=np.int64)
                for p in range(3):
                    bk.compute_ao_interleaved(p, 3, rays, off, maxd, block_samples=96)
                    acc += bk.hit_counts()
                assert np.array_equal(acc, res[key].astype(np.int64))
    assert res["simple"].sum() > 0
    assert np.array_equal(res["sample_major"], res["simple"])
    assert np.array_equal(res["stratum_major"], res["simple"])


@pytest.mark.parametrize("scale,offset,seed", [(1.0, 0.0, 1), (1e-3, 0.0, 2), (1e3, 0.0, 3), (1.0, 5e3, 4), (0.05, -2e4, 5)])
def test_fuzz_triangle_soups_match_brute_force(api, scale, offset, seed):
    """Random triangle soups (slivers, tiny and huge triangles, scenes far from the origin) and random
    rays (axis-parallel, origins on triangles): the GPU traversal equals the oracle's brute force."""
    rng = np.random.default_rng(seed)
    n, m = 3000, 40000
    c = rng.uniform(-1, 1, (n, 1, 3)) * scale
    size = (10.0 ** rng.uniform(-3, 0, (n, 1, 1))) * scale
    tri = c + rng.normal(size=(n, 3, 3)) * size
    tri[: n // 10, 2] = tri[: n // 10, 0] + (tri[: n // 10, 1] - tri[: n // 10, 0]) * rng.uniform(0.4, 0.6, (n // 10, 1)) \
        + rng.normal(size=(n // 10, 3)) * size[: n // 10, 0] * 1e-4
    tri = np.ascontiguousarray((tri + offset).reshape(n, 9), dtype=np.float32)
    rays = np.zeros((m, 8), dtype=np.float32)
    rays[:, 0:3] = rng.uniform(-1.5, 1.5, (m, 3)) * scale + offset
    d = rng.normal(size=(m, 3))
    d[: m // 20, rng.integers(0, 3)] = 0.0
    d[m // 20: m // 10] = np.eye(3)[rng.integers(0, 3, m // 10 - m // 20)] * rng.choice([-1.0, 1.0], (m // 10 - m // 20, 1))
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    k = m // 10
    pick = rng.integers(0, n, k)
    b = rng.dirichlet([1, 1, 1], k).astype(np.float32)
    rays[-k:, 0:3] = (tri[pick].reshape(k, 3, 3) * b[:, :, None]).sum(axis=1)
    rays[:, 7] = rng.uniform(0.1, 4.0, m).astype(np.float32) * scale
    mesh = Mesh(tri.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32).reshape(n, 3))
    scene = Scene([mesh], [Instance(0)])
    want = Oracle(scene).trace_rays(rays, brute=True)
    with api.Baker() as bk:
        bk.set_scene(scene)
        got = bk.trace_rays(rays)
    assert 0.01 < want.mean() < 0.99
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["sphere_ground", "warped_ground", "instanced"])
def test_oversized_primitives_hang_off_an_extra_root(api, name):
    """A ground-plane quad (two triangles spanning 100 x the scene) under a fine mesh, and the same quad as a
    huge instance in a TLAS: kept out of the tree, the hit counts are identical to the one-tree build and
    the traversal visits fewer nodes."""
    scene, blockers = SCENES[name]
    off, maxd = scenes.default_distances(scene)
    res = {}
    for split_off in (False, True):
        with api.Baker(trace_kernel=2, collect_stats=True, no_oversized_split=split_off) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(2, 0)
            bk.sample_instances(per, 2, download=False)
            bk.compute_ao(64, off, maxd, download=False)
            st = bk.stats()
            res[split_off] = (bk.hit_counts(), st.node_visits / st.rays)
    assert np.array_equal(res[False][0], res[True][0])
    assert res[False][1] < res[True][1]               # fewer node visits per ray with the extra root


def test_many_large_triangles_stay_in_the_tree(api):
    """More oversized primitives than the extra root can hold (a coarse box room around a small sphere): the
    builder falls back to one tree; brute-force parity either way."""
    rng = np.random.default_rng(21)
    quads = []
    for k in range(10):   # 20 scene-sized triangles
        a = rng.uniform(-50, 50, size=(4, 3)).astype(np.float32)
        quads.append(a)
    v = np.concatenate(quads + [scenes.uv_sphere(12, 12).vertices])
    t = np.concatenate([np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32) + 4 * k for k in range(10)] +
                       [scenes.uv_sphere(12, 12).tris + 40])
    scene = Scene([Mesh(v, t)], [Instance(0)])
    orc = Oracle(scene, Scene([], []))
    with api.Baker() as bk:
        bk.set_scene(scene)
        o = rng.uniform(-2, 2, size=(4000, 3)).astype(np.float32)
        d = rng.normal(size=(4000, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = np.concatenate([o, np.zeros((4000, 1), np.float32), d, np.full((4000, 1), 1e4, np.float32)], axis=1)
        got = bk.trace_rays(rays)
    want = orc.trace_rays(rays, brute=True)
    check_hits(orc, rays, got, want)


def test_geometry_only_scene_serves_sampling_and_vertex_maps(api):
    """aobake_set_scene_geometry (what the one-shot bake::distributeSamples / sampleInstances / mapAOToVertices
    shims use): no BVH is built, sampling and the vertex maps match a full context, tracing is a state error."""
    scene, blockers = SCENES["sphere_ground"]
    off, maxd = scenes.default_distances(scene)
    with api.Baker() as full, api.Baker() as geo:
        full.set_scene(scene, blockers)
        geo.set_scene_geometry(scene)
        assert geo.stats().num_bvh_nodes == 0
        ta, pa = full.distribute_samples(2, 0)
        tb, pb = geo.distribute_samples(2, 0)
        assert ta == tb and np.array_equal(pa, pb)
        sa = full.sample_instances(pa, 2)
        assert_samples_equal(sa, geo.sample_instances(pb, 2))
        ao = full.compute_ao(64, off, maxd)
        with pytest.raises(api.AoBakeError) as e:
            geo.compute_ao(64, off, maxd)
        assert e.value.code == 3      # AOBAKE_ERR_STATE
        geo.set_ao(ao)
        for mode in (api.FILTER_AREA_BASED, api.FILTER_LEAST_SQUARES):
            va, vb = full.map_ao_to_vertices(mode, 0.1)[0], geo.map_ao_to_vertices(mode, 0.1)[0]
            assert np.abs(va - vb).max() < 1e-5


@pytest.mark.parametrize("name", ["sphere_ground", "instanced", "sheared"])
@pytest.mark.parametrize("energy", [0, 1])
def test_least_squares_assembled_matrix_equals_matrix_free(api, name, energy):
    """The PCG product through the assembled sliced-ELL matrix (default) and the matrix-free scatter give the same
    vertex AO, and both match the oracle."""
    scene, blockers = SCENES[name]
    off, maxd = scenes.default_distances(scene)
    orc = Oracle(scene, blockers)
    out = {}
    for mf in (False, True):
        with api.Baker(ls_matrix_free=mf, ls_energy=energy, cg_tolerance=1e-9) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(2, 0)
            sb = bk.sample_instances(per, 2)
            ao = bk.compute_ao(16, off, maxd)
            out[mf] = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)
            assert bk.stats().reserved[3] == (0 if mf else 1)
    for a, b in zip(out[False], out[True]):
        assert np.abs(a - b).max() < 1e-5
    want = orc.filter_least_squares(sb, ao, 0.1, tol=1e-10, per_instance=per, energy=energy)
    for a, b in zip(out[False], want):
        assert np.abs(a - b).max() <= VERTEX_AO_TOL


def test_least_squares_high_valence_rows_are_multiplied_matrix_free(api):
    """A 48-triangle fan: the hub row has more columns than the assembly table holds; that row (and only it) is
    multiplied matrix-free — same answer as the oracle."""
    n = 48
    ang = 2 * np.pi * np.arange(n) / n
    v = np.concatenate([[[0, 0.3, 0]], np.stack([np.cos(ang), 0.05 * np.sin(3 * ang), np.sin(ang)], axis=1)]).astype(np.float32)
    t = np.array([[0, 1 + (k + 1) % n, 1 + k] for k in range(n)], dtype=np.uint32)
    scene = Scene([Mesh(v, t)], [Instance(0)])
    blockers = scenes.ground_blockers(scene)
    off, maxd = scenes.default_distances(scene)
    orc = Oracle(scene, blockers)
    with api.Baker(cg_tolerance=1e-9) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(4, 0)
        sb = bk.sample_instances(per, 4)
        ao = bk.compute_ao(16, off, maxd)
        got = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES, 0.1)[0]
        assert bk.stats().reserved[3] == 1 and bk.stats().reserved[4] == 1     # assembled matrix, one row outside it
    want = orc.filter_least_squares(sb, ao, 0.1, tol=1e-10, per_instance=per)[0]
    assert np.abs(got - want).max() <= VERTEX_AO_TOL
