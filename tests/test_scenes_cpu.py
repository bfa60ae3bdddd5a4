"""Host-side helpers (no GPU, no oracle): the procedural BASELINE.json scenes and the packing of
scenes into the C-ABI structs."""
import ctypes as C

import numpy as np

from optix_prime_baking_b200 import scenes
from optix_prime_baking_b200.ctypes_types import AoInstance, AoMesh, AoSampleInfo, AoSamples, AoScene, PackedScene, SampleBuffers


def _outward_fraction(mesh, center):
    v = mesh.vertices.astype(np.float64)
    t = mesh.tris
    n = np.cross(v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]])
    c = v[t].mean(axis=1) - center
    return ((n * c).sum(axis=1) > 0).mean()


def test_config_shapes_match_baseline_json():
    s1, b1 = scenes.config1_sphere()
    assert s1.num_triangles == 2 * 200 * 199 == 79600 and b1.num_triangles == 2           # ~80k tris + ground quad
    s2, b2 = scenes.config2_heightfield()
    assert s2.num_triangles == 2 * 708 * 708 == 1002528 and len(b2.instances) == 0         # ~1M tris
    m4 = scenes.uv_sphere(158, 158, displace=0.15, seed=4)
    assert len(m4.tris) == 2 * 158 * 157 == 49612                                           # ~50k-tri instance mesh
    s4, _ = scenes.config4_instanced(grid=3, stacks=12, slices=12)
    assert len(s4.instances) == 27 and len(s4.meshes) == 1
    # rotations are orthonormal, translations on the 2.5 x extent lattice
    for inst in s4.instances:
        r = inst.xform[:3, :3].astype(np.float64)
        assert np.allclose(r @ r.T, np.eye(3), atol=1e-5) and abs(np.linalg.det(r) - 1) < 1e-5


def test_meshes_are_consistently_wound_and_normalised():
    sph = scenes.uv_sphere(24, 24)
    assert _outward_fraction(sph, np.zeros(3)) == 1.0
    assert np.allclose(np.linalg.norm(sph.normals, axis=1), 1.0, atol=1e-6)
    hf = scenes.heightfield(32, seed=1)
    v = hf.vertices.astype(np.float64)
    n = np.cross(v[hf.tris[:, 1]] - v[hf.tris[:, 0]], v[hf.tris[:, 2]] - v[hf.tris[:, 0]])
    assert (n[:, 1] > 0).all()                                                              # terrain faces +Y
    warped = scenes.heightfield(32, seed=3, warp=0.6)
    areas = 0.5 * np.linalg.norm(np.cross(warped.vertices[warped.tris[:, 1]] - warped.vertices[warped.tris[:, 0]],
                                          warped.vertices[warped.tris[:, 2]] - warped.vertices[warped.tris[:, 0]]), axis=1)
    assert areas.max() / areas.min() > 4.0                                                  # config 3 needs non-uniform areas


def test_scenes_are_deterministic():
    a, _ = scenes.config2_heightfield(48, seed=1)
    b, _ = scenes.config2_heightfield(48, seed=1)
    c, _ = scenes.config2_heightfield(48, seed=2)
    assert np.array_equal(a.meshes[0].vertices, b.meshes[0].vertices)
    assert not np.array_equal(a.meshes[0].vertices, c.meshes[0].vertices)


def test_default_distances_and_ground_follow_the_cli_defaults():
    scene, blockers = scenes.config1_sphere(16, 16)
    off, maxd = scenes.default_distances(scene)
    assert abs(off - 0.02) < 1e-6 and abs(maxd - 20.0) < 1e-4                                # 0.01 / 10 x extent 2
    g = blockers.meshes[0]
    assert np.allclose(g.vertices[:, 1], -1.0 - 0.03 * 2.0) and abs(np.abs(g.vertices[:, 0]).max() - 100.0) < 1e-3


def test_struct_layouts_match_the_c_header():
    # sizes/offsets the C compiler produces for include/aobake.h on x86-64
    assert C.sizeof(AoSampleInfo) == 20
    assert C.sizeof(AoMesh) == 80 and AoMesh.num_triangles.offset == 40 and AoMesh.bbox_min.offset == 56
    assert C.sizeof(AoInstance) == 104 and AoInstance.storage_identifier.offset == 64 and AoInstance.mesh_index.offset == 72
    assert C.sizeof(AoScene) == 32 and C.sizeof(AoSamples) == 40


def test_packed_scene_points_at_the_numpy_buffers():
    scene, _ = scenes.config4_instanced(grid=2, stacks=8, slices=8)
    ps = PackedScene(scene)
    assert ps.c.num_meshes == 1 and ps.c.num_instances == 8
    m = ps.meshes[0]
    assert m.vertices == scene.meshes[0].vertices.ctypes.data and m.vertex_stride_bytes == 12
    assert m.num_triangles == len(scene.meshes[0].tris)
    assert np.allclose(np.array(ps.instances[5].xform[:]).reshape(4, 4), scene.instances[5].xform)
    sb = SampleBuffers(10)
    assert sb.c.num_samples == 10 and sb.infos.dtype.itemsize == 20
