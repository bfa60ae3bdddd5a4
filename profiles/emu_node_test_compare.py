#!/usr/bin/env python
"""CPU emulation (tests/emu): node visits and triangle tests per AO ray with the packed-fp16 node
test against the fp32 one, on scaled-down versions of the bench workloads.  The hit results must be
identical (both tests are conservative; the triangle test decides).  No GPU needed.
usage: emu_node_test_compare.py [rays_per_sample]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optix_prime_baking_b200 import scenes  # noqa: E402
from tests.oracle_binding import Oracle  # noqa: E402
from tests import oracle_binding as ob  # noqa: E402

subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
E_ = C.CDLL(os.path.join(ROOT, "tests", "emu", "libaob_emu.so"))
E_.emu_bvh_create_flat.restype = C.c_void_p
E_.emu_bvh_create_flat.argtypes = [C.c_void_p, C.c_uint32]
E_.emu_bvh_create_two_level.restype = C.c_void_p
E_.emu_bvh_create_two_level.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
E_.emu_trace.restype = C.c_uint64
E_.emu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
E_.emu_bvh_destroy.argtypes = [C.c_void_p]
E_.emu_set_max_leaf.argtypes = [C.c_uint32]
E_.emu_set_node_test_fp32.argtypes = [C.c_int]

rps = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def world_tris(sc_list):
    out = []
    for sc in sc_list:
        if sc is None:
            continue
        for inst in sc.instances:
            m = sc.meshes[inst.mesh_index]
            w = (m.vertices.astype(np.float64) @ inst.xform[:3, :3].T.astype(np.float64) + inst.xform[:3, 3].astype(np.float64)).astype(np.float32)
            out.append(w[m.tris].reshape(-1, 9))
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)


def ao_rays(scene, blk, n_samples):
    orc = Oracle(scene, blk)
    _, per = orc.distribute_samples(0, n_samples)
    sb = orc.sample_instances(per, 0)
    off, md = scenes.default_distances(scene)
    return np.ascontiguousarray(orc.generate_rays(sb, 0, sb.n, rps, off, md).reshape(-1, 8), dtype=np.float32)


def run(B, rays):
    res = []
    for fp32 in (1, 0):
        E_.emu_set_node_test_fp32(fp32)
        hit = np.zeros(len(rays), dtype=np.uint8)
        tt = C.c_uint64()
        nodes = E_.emu_trace(B, rays.ctypes.data, len(rays), hit.ctypes.data, C.byref(tt))
        res.append((nodes / len(rays), tt.value / len(rays), hit))
    assert np.array_equal(res[0][2], res[1][2]), "fp16 and fp32 node tests disagree on hits"
    return res


def report(name, res):
    (n32, t32, h), (n16, t16, _) = res
    print(f"{name:34s} hit {h.mean():.3f} | nodes/ray fp32 {n32:.3f} fp16 {n16:.3f} ({100 * (n16 / n32 - 1):+.2f} %) | tris/ray fp32 {t32:.3f} fp16 {t16:.3f} ({100 * (t16 / max(t32, 1e-9) - 1):+.2f} %)")


E_.emu_set_max_leaf(2)
for name, (scene, blk) in {"c2-like heightfield 256^2": scenes.config2_heightfield(256), "c1 sphere 100x100 + ground": scenes.config1_sphere(100, 100),
                           "c3-like warped terrain 300^2": scenes.config3_bigmesh(300)}.items():
    rays = ao_rays(scene, blk, 20000)
    wt = world_tris([scene, blk])
    B = E_.emu_bvh_create_flat(wt.ctypes.data, len(wt))
    report(name, run(B, rays))
    E_.emu_bvh_destroy(B)

E_.emu_set_max_leaf(1)
scene, blk = scenes.config4_instanced(grid=5, stacks=40, slices=40, with_ground=False)
meshes = list(scene.meshes)
soups = [np.ascontiguousarray(m.vertices[m.tris].reshape(-1, 9), dtype=np.float32) for m in meshes]
ptrs = (C.c_void_p * len(soups))(*[s.ctypes.data for s in soups])
ntris = np.array([len(s) for s in soups], dtype=np.uint32)
imesh = np.array([i.mesh_index for i in scene.instances], dtype=np.uint32)
xf = np.ascontiguousarray(np.stack([i.xform for i in scene.instances]), dtype=np.float32).reshape(-1, 16)
inv = np.zeros((len(imesh), 12), dtype=np.float32)
for k in range(len(imesh)):
    ob.lib().ao_oracle_affine_inverse(xf[k].ctypes.data, inv[k].ctypes.data)
B = E_.emu_bvh_create_two_level(len(soups), ptrs, ntris.ctypes.data, len(imesh), imesh.ctypes.data, xf.ctypes.data, inv.ctypes.data)
report("c4-like 5^3 instanced lattice (TLAS)", run(B, ao_rays(scene, None, 20000)))
E_.emu_bvh_destroy(B)
