"""CPU tests of the oracle (the checker): golden vectors, analytic known answers, brute-force
cross-checks, and the host-emulated kernel bodies against it.  No GPU needed."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from optix_prime_baking_b200 import scenes
from optix_prime_baking_b200.scenes import Instance, Mesh, Scene

from . import oracle_binding as ob
from .oracle_binding import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "oracle_golden.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu")], check=True)
    E = C.CDLL(os.path.join(HERE, "emu", "libaob_emu.so"))
    E.emu_bvh_create_flat.restype = C.c_void_p
    E.emu_bvh_create_flat.argtypes = [C.c_void_p, C.c_uint32]
    E.emu_bvh_create_two_level.restype = C.c_void_p
    E.emu_bvh_create_two_level.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    E.emu_bvh_destroy.argtypes = [C.c_void_p]
    E.emu_trace.restype = C.c_uint64
    E.emu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    E.emu_generate_rays.argtypes = [C.c_void_p] * 3 + [C.c_uint64, C.c_uint64, C.c_int, C.c_float, C.c_float, C.c_void_p]
    E.emu_tea.restype = C.c_uint32
    E.emu_tea.argtypes = [C.c_uint32] * 3
    E.emu_halton.restype = C.c_float
    E.emu_halton.argtypes = [C.c_uint32] * 2
    E.emu_sincos2pi.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    return E


def test_rng_golden(golden):
    for rounds, v0, v1, want in golden["tea"]:
        assert ob.tea(rounds, v0, v1) == int(want, 16)
        assert int(scenes.tea(rounds, np.uint32(v0), np.uint32(v1))) == int(want, 16)
    for seed, stream in golden["lcg"]:
        assert ob.lcg_stream(seed, len(stream)) == [int(x, 16) for x in stream]
    for i, base, want in golden["halton"]:
        assert np.float32(ob.halton(i, base)).view(np.uint32) == int(want, 16)


def test_whole_bake_golden():
    """A complete small bake (samples, rays, hit counts, vertex AO) against the committed fixture
    tests/golden/bake_golden.npz (written by tests/golden/make_golden.py from the oracle)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bake_golden.npz"))
    scene, blockers = scenes.config1_sphere(8, 10)
    orc = Oracle(scene, blockers)
    total, per = orc.distribute_samples(2, 0)
    sb = orc.sample_instances(per, 2)
    off, maxd = scenes.default_distances(scene)
    assert np.float32(off) == g["offset"] and np.float32(maxd) == g["maxdist"]
    assert np.array_equal(sb.infos["tri_idx"], g["tri_idx"])
    for name, arr in [("bary", sb.infos["bary"]), ("dA", sb.infos["dA"]), ("positions", sb.positions), ("normals", sb.normals),
                      ("face_normals", sb.face_normals)]:
        assert np.array_equal(np.ascontiguousarray(arr).view(np.uint32), g[name]), name
    assert np.array_equal(orc.generate_rays(sb, 0, 8, 16, off, maxd).view(np.uint32), g["rays_first8"])
    ao, hits = orc.compute_ao(sb, 16, off, maxd)
    assert np.array_equal(hits, g["hits"]) and np.array_equal(ao.view(np.uint32), g["ao"])
    assert np.array_equal(orc.filter_area(sb, ao)[0].view(np.uint32), g["v_area"])
    assert np.abs(orc.filter_least_squares(sb, ao, 0.1, tol=1e-12)[0] - g["v_ls"]).max() < 1e-6


def test_rng_properties():
    # lcg low 24 bits; rnd in [0,1); halton(1,2)=0.5, halton(2,2)=0.25, halton(1,3)=1/3
    assert all(0 <= x < (1 << 24) for x in ob.lcg_stream(12345, 1000))
    assert ob.halton(1, 2) == 0.5 and ob.halton(2, 2) == 0.25 and ob.halton(3, 2) == 0.75
    assert abs(ob.halton(1, 3) - 1 / 3) < 1e-7
    for u in np.linspace(0, 0.999, 257, dtype=np.float32):
        c, s = ob.sincos2pi(float(u))
        assert abs(c - np.cos(2 * np.pi * float(u))) < 1e-6 and abs(s - np.sin(2 * np.pi * float(u))) < 1e-6


def test_sample_budget_rules():
    scene, _ = scenes.config3_bigmesh(40, seed=3)
    orc = Oracle(scene)
    nT = scene.num_triangles
    for mn, req in [(3, 0), (0, 1000), (1, 10007), (2, 1)]:
        total, per = orc.distribute_samples(mn, req)
        assert total == max(req, mn * nT) and per.sum() == total
        counts = orc.triangle_counts(0, total, mn)
        assert counts.sum() == total and counts.min() >= mn
        sb = orc.sample_instances(per, mn)
        assert np.array_equal(np.bincount(sb.infos["tri_idx"], minlength=nT), counts)
        b = sb.infos["bary"]
        assert (b >= -1e-7).all() and np.allclose(b.sum(axis=1), 1.0, atol=1e-6)
        # sum of dA over a triangle = its area
        area = np.bincount(sb.infos["tri_idx"], weights=sb.infos["dA"].astype(np.float64), minlength=nT)
        m = scene.meshes[0]
        v = m.vertices.astype(np.float64)
        ta = 0.5 * np.linalg.norm(np.cross(v[m.tris[:, 1]] - v[m.tris[:, 0]], v[m.tris[:, 2]] - v[m.tris[:, 0]]), axis=1)
        assert np.allclose(area[counts > 0], ta[counts > 0], rtol=1e-5)
    # leftover rule with equal areas and < 1 sample per triangle: the first N triangles get one each
    flat = Scene([scenes.heightfield(8, seed=1, height=0.0)], [Instance(0)])
    o2 = Oracle(flat)
    counts = o2.triangle_counts(0, 50, 0)
    assert counts.sum() == 50 and set(counts.tolist()) <= {0, 1}


def test_cosine_law_and_hemisphere():
    scene, blk = scenes.config1_sphere(24, 24)
    orc = Oracle(scene, blk)
    _, per = orc.distribute_samples(1, 0)
    sb = orc.sample_instances(per, 1)
    rays = orc.generate_rays(sb, 0, sb.n, 256, 0.02, 20.0)
    d = rays[:, :, 4:7]
    assert np.abs(np.linalg.norm(d, axis=2) - 1).max() < 1e-5
    cosn = (d * sb.normals[:, None, :]).sum(axis=2)
    assert abs(cosn.mean() - 2.0 / 3.0) < 2e-3          # E[cos] = 2/3 under cosine sampling
    assert ((d * sb.face_normals[:, None, :]).sum(axis=2) > 0).mean() > 0.999
    o = rays[:, 0, 0:3]
    assert np.allclose(o, sb.positions + np.float32(0.02) * sb.normals, atol=1e-6)


def test_analytic_ao_sphere_over_plane():
    scene, _ = scenes.config1_sphere(40, 40)
    blockers = scenes.ground_blockers(scene, 1, 1.0e6, 0.03)
    orc = Oracle(scene, blockers)
    _, per = orc.distribute_samples(3, 0)
    sb = orc.sample_instances(per, 3)
    ao, _ = orc.compute_ao(sb, 1024, 0.02, 1.0e7)
    expect = (1.0 + sb.normals[:, 1]) / 2.0
    assert np.abs(ao - expect).mean() < 5e-3 and np.abs(ao - expect).max() < 6e-2


def _random_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene.world_bbox()
    ext = float((hi - lo).max())
    r = np.zeros((n, 8), dtype=np.float32)
    r[:, 0:3] = rng.uniform(lo - 0.2 * ext, hi + 0.2 * ext, (n, 3))
    d = rng.normal(size=(n, 3))
    r[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    r[:, 7] = rng.uniform(0.05 * ext, 3 * ext, n)
    return r


@pytest.mark.parametrize("mode", [1, 2])
def test_oracle_bvh_vs_brute_force(mode):
    scene, blk = scenes.config4_instanced(grid=2, stacks=10, slices=10, with_ground=True)
    orc = Oracle(scene, blk, mode)
    rays = _random_rays(scene, 20000, 3)
    b = orc.trace_rays(rays, brute=True)
    kinds = set()
    try:   # both walks of the oracle (binary scalar, 8-wide AVX2 where the CPU has it) against brute force
        for trav in (ob.TRAVERSAL_BINARY, ob.TRAVERSAL_WIDE):
            kinds.add(ob.set_traversal(trav))
            a = orc.trace_rays(rays)
            assert np.array_equal(a, b) and 0.02 < a.mean() < 0.98
    finally:
        ob.set_traversal(ob.TRAVERSAL_AUTO)
    assert 1 in kinds


def _world_tris(scene_list):
    out = []
    for sc in scene_list:
        for inst in sc.instances:
            m = sc.meshes[inst.mesh_index]
            v, xf = m.vertices, inst.xform.astype(np.float32)
            w = np.empty_like(v)
            for r in range(3):
                w[:, r] = ((xf[r, 0] * v[:, 0] + xf[r, 1] * v[:, 1]) + xf[r, 2] * v[:, 2]) + xf[r, 3]
            out.append(w[m.tris].reshape(-1, 9))
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)


def test_emulated_kernels_match_oracle(emu):
    """The LBVH build, 8-wide collapse, quantised traversal and ray generation that the CUDA
    kernels run, executed serially on the CPU, agree with the oracle."""
    for rounds, v0, v1 in [(2, 1, 2), (4, 0, 0), (4, 123, 456789)]:
        assert emu.emu_tea(rounds, v0, v1) == ob.tea(rounds, v0, v1)
    for i in range(1, 200):
        assert emu.emu_halton(i, 2) == ob.halton(i, 2) and emu.emu_halton(i, 3) == ob.halton(i, 3)
    for name, (scene, blk) in {"sphere": scenes.config1_sphere(30, 30), "hf": scenes.config2_heightfield(40)}.items():
        orc = Oracle(scene, blk)
        _, per = orc.distribute_samples(1, 0)
        sb = orc.sample_instances(per, 1)
        off, md = scenes.default_distances(scene)
        n = min(sb.n, 1500)
        want = orc.generate_rays(sb, 0, n, 16, off, md)
        got = np.zeros_like(want)
        emu.emu_generate_rays(sb.positions.ctypes.data, sb.normals.ctypes.data, sb.face_normals.ctypes.data, 0, n, 16,
                              float(off), float(md), got.ctypes.data)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
        rays = np.ascontiguousarray(np.concatenate([want.reshape(-1, 8), _random_rays(scene, 8000, 1)]), dtype=np.float32)
        wt = _world_tris([scene, blk])
        B = emu.emu_bvh_create_flat(wt.ctypes.data, len(wt))
        hit = np.zeros(len(rays), dtype=np.uint8)
        emu.emu_trace(B, rays.ctypes.data, len(rays), hit.ctypes.data, None)
        emu.emu_bvh_destroy(B)
        assert np.array_equal(hit, orc.trace_rays(rays)), name


def test_emulated_oversized_primitives_on_an_extra_root(emu):
    """The builder's split of oversized primitives (the ground-plane quad under a fine mesh: aob_kernels.cuh
    k_flag_big / k_super_root, emulated through the shared bodies): identical hit/miss decisions with and without
    the split, both equal to the oracle, and fewer node visits with it."""
    sc, _ = scenes.config3_bigmesh(60, seed=3)
    blk = scenes.ground_blockers(sc)
    orc = Oracle(sc, blk)
    _, per = orc.distribute_samples(1, 0)
    sb = orc.sample_instances(per, 1)
    off, md = scenes.default_distances(sc)
    rays = np.ascontiguousarray(np.concatenate([orc.generate_rays(sb, 0, min(sb.n, 2000), 16, off, md).reshape(-1, 8), _random_rays(sc, 4000, 3)]),
                                dtype=np.float32)
    wt = _world_tris([sc, blk])
    want = orc.trace_rays(rays)
    visits = {}
    for off_switch in (0, 1):
        emu.emu_set_no_oversized_split(off_switch)
        B = emu.emu_bvh_create_flat(wt.ctypes.data, len(wt))
        hit = np.zeros(len(rays), dtype=np.uint8)
        emu.emu_trace.restype = C.c_uint64
        visits[off_switch] = emu.emu_trace(B, rays.ctypes.data, len(rays), hit.ctypes.data, None)
        emu.emu_bvh_destroy(B)
        assert np.array_equal(hit, want), off_switch
    emu.emu_set_no_oversized_split(0)
    assert visits[0] < 0.8 * visits[1]      # the extra root keeps the coarse grid out of the tree


def test_emulated_two_level_matches_oracle(emu):
    scene, blk = scenes.config4_instanced(grid=2, stacks=10, slices=10, with_ground=True)
    orc = Oracle(scene, blk, 2)
    meshes = [m for m in scene.meshes] + [m for m in blk.meshes]
    soups = [np.ascontiguousarray(m.vertices[m.tris].reshape(-1, 9), dtype=np.float32) for m in meshes]
    ptrs = (C.c_void_p * len(soups))(*[s.ctypes.data for s in soups])
    ntris = np.array([len(s) for s in soups], dtype=np.uint32)
    insts = [(i.mesh_index, i.xform) for i in scene.instances] + [(len(scene.meshes) + i.mesh_index, i.xform) for i in blk.instances]
    imesh = np.array([m for m, _ in insts], dtype=np.uint32)
    xf = np.ascontiguousarray(np.stack([x for _, x in insts]), dtype=np.float32).reshape(-1, 16)
    inv = np.zeros((len(insts), 12), dtype=np.float32)
    for k in range(len(insts)):
        ob.lib().ao_oracle_affine_inverse(xf[k].ctypes.data, inv[k].ctypes.data)
    B = emu.emu_bvh_create_two_level(len(soups), ptrs, ntris.ctypes.data, len(insts), imesh.ctypes.data, xf.ctypes.data, inv.ctypes.data)
    rays = _random_rays(scene, 20000, 5)
    hit = np.zeros(len(rays), dtype=np.uint8)
    emu.emu_trace(B, rays.ctypes.data, len(rays), hit.ctypes.data, None)
    emu.emu_bvh_destroy(B)
    want = orc.trace_rays(rays)
    assert (hit != want).mean() <= 1e-4 and 0.02 < want.mean() < 0.98


def test_area_filter_properties():
    scene, blk = scenes.config1_sphere(16, 16)
    orc = Oracle(scene, blk)
    _, per = orc.distribute_samples(4, 0)
    sb = orc.sample_instances(per, 4)
    const = orc.filter_area(sb, np.full(sb.n, 0.37, dtype=np.float32))[0]
    assert np.allclose(const, 0.37, atol=1e-6)           # partition of unity
    ao, _ = orc.compute_ao(sb, 64, 0.02, 20.0)
    v = orc.filter_area(sb, ao)[0]
    assert v.min() >= ao.min() - 1e-6 and v.max() <= ao.max() + 1e-6


def test_least_squares_filter_algebra():
    """w = 0 and AO exactly linear in the barycentrics returns the generating vertex values; a
    linear field on a planar mesh is in the regulariser's null space."""
    flat = Scene([scenes.heightfield(12, seed=2, height=0.0)], [Instance(0)])
    m = flat.meshes[0]
    orc = Oracle(flat)
    _, per = orc.distribute_samples(6, 0)
    sb = orc.sample_instances(per, 6)
    rng = np.random.default_rng(0)
    xv = rng.uniform(0, 1, len(m.vertices)).astype(np.float32)
    tri = m.tris[sb.infos["tri_idx"]]
    ao = (sb.infos["bary"] * xv[tri]).sum(axis=1).astype(np.float32)
    got = orc.filter_least_squares(sb, ao, weight=0.0, tol=1e-12)[0]
    assert np.abs(got - xv).max() < 1e-4
    lin = (0.2 + 0.03 * m.vertices[:, 0] - 0.05 * m.vertices[:, 2]).astype(np.float32)
    ao_lin = (sb.infos["bary"] * lin[tri]).sum(axis=1).astype(np.float32)
    got = orc.filter_least_squares(sb, ao_lin, weight=10.0, tol=1e-12)[0]
    assert np.abs(got - lin).max() < 1e-4                # R x = 0 for linear x, so the fit is exact
    # regularisation pulls a noisy field towards smoothness but keeps its mean
    noisy = (0.5 + 0.2 * rng.standard_normal(sb.n)).astype(np.float32)
    a, b = orc.filter_least_squares(sb, noisy, 0.0)[0], orc.filter_least_squares(sb, noisy, 1.0)[0]
    assert np.var(b) < np.var(a) and abs(a.mean() - b.mean()) < 0.02


def test_ground_plane_matches_host_helper():
    lo, hi = np.array([-1, -1, -1], dtype=np.float32), np.array([1, 2, 3], dtype=np.float32)
    for up in range(6):
        v = np.zeros((4, 3), dtype=np.float32)
        t = np.zeros((2, 3), dtype=np.uint32)
        ob.lib().ao_oracle_make_ground_plane(lo.ctypes.data, hi.ctypes.data, up, 100.0, 0.03, v.ctypes.data, t.ctypes.data)
        g = scenes.ground_plane(lo, hi, up, 100.0, 0.03)
        assert np.allclose(v, g.vertices) and np.array_equal(t, g.tris)
        n = np.cross(v[t[0, 1]] - v[t[0, 0]], v[t[0, 2]] - v[t[0, 0]])
        assert (n[up % 3] > 0) == (up < 3)               # faces the scene


def test_woop_variants_agree_bit_for_bit(emu):
    """The axis-specialised and the select-based Woop tests (chosen per warp on the GPU) make
    identical decisions for every (ray, triangle) pair."""
    emu.emu_woop_variants_agree.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    sc, bl = scenes.config1_sphere(12, 12)
    wt = _world_tris([sc, bl])
    rays = _random_rays(sc, 2000, 9)
    assert emu.emu_woop_variants_agree(rays.ctypes.data, len(rays), wt.ctypes.data, len(wt)) == 0


@pytest.mark.parametrize("scale,offset,seed", [(1.0, 0.0, 1), (1e-3, 0.0, 2), (1e3, 0.0, 3), (1.0, 5e3, 4), (0.05, -2e4, 5), (1.0, 0.0, 6)])
def test_fuzz_quantised_traversal_is_conservative(emu, scale, offset, seed):
    """Random triangle soups (sliver, tiny and huge triangles, scenes far from the origin) and random
    rays (including axis-parallel ones and origins on triangle planes): the emulated 8-wide quantised
    traversal must report exactly the brute-force any-hit answer — a box culled by rounding would
    show up as a false miss."""
    rng = np.random.default_rng(seed)
    n = 3000
    c = rng.uniform(-1, 1, (n, 1, 3)) * scale
    size = (10.0 ** rng.uniform(-3, 0, (n, 1, 1))) * scale
    tri = c + rng.normal(size=(n, 3, 3)) * size
    tri[: n // 10, 2] = tri[: n // 10, 0] + (tri[: n // 10, 1] - tri[: n // 10, 0]) * rng.uniform(0.4, 0.6, (n // 10, 1)) \
        + rng.normal(size=(n // 10, 3)) * size[: n // 10, 0] * 1e-4                      # slivers
    tri = np.ascontiguousarray((tri + offset).reshape(n, 9), dtype=np.float32)
    m = 20000
    rays = np.zeros((m, 8), dtype=np.float32)
    rays[:, 0:3] = rng.uniform(-1.5, 1.5, (m, 3)) * scale + offset
    d = rng.normal(size=(m, 3))
    d[: m // 20, rng.integers(0, 3)] = 0.0                                                 # axis-parallel components
    d[m // 20: m // 10] = np.eye(3)[rng.integers(0, 3, m // 10 - m // 20)] * rng.choice([-1.0, 1.0], (m // 10 - m // 20, 1))
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    # a tenth of the rays start on a triangle (t = 0 exactly at the origin's own triangle)
    k = m // 10
    pick = rng.integers(0, n, k)
    b = rng.dirichlet([1, 1, 1], k).astype(np.float32)
    rays[-k:, 0:3] = (tri[pick].reshape(k, 3, 3) * b[:, :, None]).sum(axis=1)
    rays[:, 7] = rng.uniform(0.1, 4.0, m).astype(np.float32) * scale
    from optix_prime_baking_b200.scenes import Instance, Mesh, Scene
    mesh = Mesh(tri.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32).reshape(n, 3))
    orc = Oracle(Scene([mesh], [Instance(0)]))
    want = orc.trace_rays(rays, brute=True)
    try:                                                                                   # the oracle's own BVH, both walks
        for trav in (ob.TRAVERSAL_BINARY, ob.TRAVERSAL_WIDE):
            ob.set_traversal(trav)
            assert np.array_equal(orc.trace_rays(rays), want), trav
    finally:
        ob.set_traversal(ob.TRAVERSAL_AUTO)
    B = emu.emu_bvh_create_flat(tri.ctypes.data, n)
    emu.emu_trace.restype = C.c_uint64
    visits = []
    for fp32_only in (0, 1):        # 0: packed-fp16 node test (fp32 for the rays outside its range); 1: fp32 for every ray
        emu.emu_set_node_test_fp32(fp32_only)
        hit = np.zeros(m, dtype=np.uint8)
        visits.append(emu.emu_trace(B, rays.ctypes.data, m, hit.ctypes.data, None))
        assert np.array_equal(hit, want), fp32_only
    emu.emu_set_node_test_fp32(0)
    emu.emu_bvh_destroy(B)
    assert 0.01 < want.mean() < 0.99
    assert 0.99 * visits[1] <= visits[0] <= 1.25 * visits[1]      # the fp16 boxes are a little fatter (wider padding)
