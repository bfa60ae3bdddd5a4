"""ctypes mirrors of the POD structs in include/aobake.h (bake::Mesh / Instance / Scene /
SampleInfo / AOSamples of the reference's bake_api.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .scenes import Scene


class AoMesh(C.Structure):
    _fields_ = [("num_vertices", C.c_uint64), ("vertices", C.c_void_p), ("vertex_stride_bytes", C.c_uint32),
                ("normals", C.c_void_p), ("normal_stride_bytes", C.c_uint32), ("num_triangles", C.c_uint64),
                ("tri_vertex_indices", C.c_void_p), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3)]


class AoInstance(C.Structure):
    _fields_ = [("xform", C.c_float * 16), ("storage_identifier", C.c_uint64), ("mesh_index", C.c_uint32),
                ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3)]


class AoScene(C.Structure):
    _fields_ = [("meshes", C.POINTER(AoMesh)), ("num_meshes", C.c_uint64),
                ("instances", C.POINTER(AoInstance)), ("num_instances", C.c_uint64)]


class AoSampleInfo(C.Structure):
    _fields_ = [("tri_idx", C.c_uint32), ("bary", C.c_float * 3), ("dA", C.c_float)]


class AoSamples(C.Structure):
    _fields_ = [("num_samples", C.c_uint64), ("sample_positions", C.c_void_p), ("sample_normals", C.c_void_p),
                ("sample_face_normals", C.c_void_p), ("sample_infos", C.c_void_p)]


SAMPLE_INFO_DTYPE = np.dtype([("tri_idx", np.uint32), ("bary", np.float32, (3,)), ("dA", np.float32)])
assert SAMPLE_INFO_DTYPE.itemsize == 20 == C.sizeof(AoSampleInfo)


class PackedScene:
    """Owns the ctypes arrays (and keeps the numpy buffers alive) behind an AoScene."""

    def __init__(self, scene: Scene):
        self.scene = scene
        nm, ni = len(scene.meshes), len(scene.instances)
        self.meshes = (AoMesh * max(nm, 1))()
        self.instances = (AoInstance * max(ni, 1))()
        for k, m in enumerate(scene.meshes):
            cm = self.meshes[k]
            cm.num_vertices = len(m.vertices)
            cm.vertices = m.vertices.ctypes.data
            cm.vertex_stride_bytes = m.vertices.strides[0]
            cm.normals = m.normals.ctypes.data if m.normals is not None else None
            cm.normal_stride_bytes = m.normals.strides[0] if m.normals is not None else 0
            cm.num_triangles = len(m.tris)
            cm.tri_vertex_indices = m.tris.ctypes.data
            lo, hi = m.bbox
            cm.bbox_min[:] = [float(x) for x in lo]
            cm.bbox_max[:] = [float(x) for x in hi]
        for k, inst in enumerate(scene.instances):
            ci = self.instances[k]
            xf = np.ascontiguousarray(inst.xform, dtype=np.float32).reshape(16)
            ci.xform[:] = [float(x) for x in xf]
            ci.storage_identifier = inst.storage_identifier
            ci.mesh_index = inst.mesh_index
        self.c = AoScene(self.meshes, nm, self.instances, ni)

    def ref(self):
        return C.byref(self.c)


class SampleBuffers:
    """Host SoA sample arrays (allocate_ao_samples of bake_util.cpp) + the AoSamples view."""

    def __init__(self, n: int):
        self.n = int(n)
        self.positions = np.zeros((self.n, 3), dtype=np.float32)
        self.normals = np.zeros((self.n, 3), dtype=np.float32)
        self.face_normals = np.zeros((self.n, 3), dtype=np.float32)
        self.infos = np.zeros(self.n, dtype=SAMPLE_INFO_DTYPE)
        self.c = AoSamples(self.n, self.positions.ctypes.data, self.normals.ctypes.data,
                           self.face_normals.ctypes.data, self.infos.ctypes.data)

    def ref(self):
        return C.byref(self.c)
