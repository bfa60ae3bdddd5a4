"""Python host-side mirror of the reference's bake API over the C-ABI (ctypes).

`distribute_samples / sample_instances / compute_ao / map_ao_to_vertices` are the
one-shot, host-buffer forms of bake::distributeSamples / sampleInstances / computeAO /
mapAOToVertices (bake_api.h).  `Baker` is the resident form (scene, BVH,
samples and AO stay in HBM between calls).  Everything computes in libaobake.so's CUDA
kernels; if the library is missing or no GPU is present this module raises — there is no
CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .ctypes_types import PackedScene, SampleBuffers
from .scenes import Scene

FILTER_AREA_BASED = 0
FILTER_LEAST_SQUARES = 1
INSTANCING_AUTO, INSTANCING_FLATTEN, INSTANCING_TWO_LEVEL = 0, 1, 2

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libaobake.so")
_LIB = None

EXPORTS = [
    "aobake_default_params", "aobake_create", "aobake_destroy", "aobake_last_error", "aobake_set_stream",
    "aobake_synchronize", "aobake_set_scene", "aobake_distribute_samples", "aobake_sample_instances",
    "aobake_set_samples", "aobake_compute_ao", "aobake_compute_ao_range", "aobake_compute_ao_interleaved", "aobake_comm_unique_id",
    "aobake_comm_init", "aobake_comm_destroy", "aobake_compute_ao_distributed", "aobake_map_ao_to_vertices_distributed",
    "aobake_set_scene_distributed", "aobake_set_samples_distributed", "aobake_set_scene_geometry", "aobake_get_ao_device", "aobake_set_ao",
    "aobake_map_ao_to_vertices", "aobake_make_ground_plane", "aobake_trace_rays", "aobake_dump_rays",
    "aobake_get_hit_counts", "aobake_get_timings", "aobake_get_stats", "aobake_num_samples",
]


class AoBakeParams(C.Structure):
    _fields_ = [("device", C.c_int32), ("instancing_mode", C.c_int32), ("cg_max_iterations", C.c_int32),
                ("cg_tolerance", C.c_float), ("trace_kernel", C.c_int32), ("collect_stats", C.c_int32),
                ("refill_below", C.c_int32), ("leaf_tris", C.c_int32), ("node_test", C.c_int32), ("deferred_capacity", C.c_int32), ("tri_batch", C.c_int32), ("no_oversized_split", C.c_int32), ("ls_energy", C.c_int32), ("ls_matrix_free", C.c_int32), ("ray_order", C.c_int32)]


class AoTimings(C.Structure):
    _fields_ = [("upload_ms", C.c_float), ("bvh_build_ms", C.c_float), ("sample_ms", C.c_float),
                ("trace_ms", C.c_float), ("filter_ms", C.c_float), ("host_total_ms", C.c_float),
                ("rays_traced", C.c_uint64), ("cg_iterations", C.c_int32), ("kernel_launches", C.c_int32),
                ("reserved", C.c_int32 * 6)]


class AoStats(C.Structure):
    _fields_ = [("num_bvh_nodes", C.c_uint64), ("num_bvh_triangles", C.c_uint64), ("num_tlas_instances", C.c_uint64),
                ("bvh_bytes", C.c_uint64), ("node_visits", C.c_uint64), ("triangle_tests", C.c_uint64),
                ("instance_entries", C.c_uint64), ("rays", C.c_uint64), ("two_level", C.c_int32),
                ("reserved", C.c_int32 * 7)]


class AoBakeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"aobake status {code}: {msg}")
        self.code = code


def load_library(path: Optional[str] = None):
    """Loads libaobake.so.  Raises if it has not been built: the product never falls back."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or os.environ.get("AOBAKE_LIB") or _LIB_PATH   # AOBAKE_LIB: A/B-test another build of the same ABI
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `python -m optix_prime_baking_b200.build` "
                           "(requires nvcc; there is no CPU fallback)")
    L = C.CDLL(p)
    vp, u64, sz, i32, f32 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_int, C.c_float
    L.aobake_default_params.argtypes = [vp]
    L.aobake_create.argtypes = [vp, C.POINTER(vp)]
    L.aobake_destroy.argtypes = [vp]
    L.aobake_destroy.restype = None
    L.aobake_last_error.argtypes = [vp]
    L.aobake_last_error.restype = C.c_char_p
    L.aobake_set_stream.argtypes = [vp, vp]
    L.aobake_synchronize.argtypes = [vp]
    L.aobake_set_scene.argtypes = [vp, vp, vp]
    if hasattr(L, "aobake_set_scene_geometry"):
        L.aobake_set_scene_geometry.argtypes = [vp, vp]
    if hasattr(L, "aobake_set_scene_distributed"):   # (absent only from older builds loaded through AOBAKE_LIB for A/B timing)
        L.aobake_set_scene_distributed.argtypes = [vp, vp, vp]
        L.aobake_set_samples_distributed.argtypes = [vp, vp, vp]
    L.aobake_distribute_samples.argtypes = [vp, sz, sz, vp, C.POINTER(sz)]
    L.aobake_sample_instances.argtypes = [vp, vp, sz, vp]
    L.aobake_set_samples.argtypes = [vp, vp, vp]
    L.aobake_compute_ao.argtypes = [vp, i32, f32, f32, vp]
    L.aobake_compute_ao_range.argtypes = [vp, sz, sz, i32, f32, f32, vp]
    L.aobake_compute_ao_interleaved.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, i32, f32, f32]
    L.aobake_comm_unique_id.argtypes = [vp]
    L.aobake_comm_init.argtypes = [vp, i32, i32, vp]
    L.aobake_comm_destroy.argtypes = [vp]
    L.aobake_compute_ao_distributed.argtypes = [vp, i32, f32, f32, vp]
    L.aobake_get_ao_device.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.aobake_set_ao.argtypes = [vp, vp]
    L.aobake_map_ao_to_vertices.argtypes = [vp, i32, f32, vp]
    L.aobake_map_ao_to_vertices_distributed.argtypes = [vp, i32, f32, vp]
    L.aobake_make_ground_plane.argtypes = [vp, vp, i32, f32, f32, vp, vp]
    L.aobake_trace_rays.argtypes = [vp, vp, sz, vp]
    L.aobake_dump_rays.argtypes = [vp, sz, sz, i32, f32, f32, vp]
    L.aobake_get_hit_counts.argtypes = [vp, vp]
    L.aobake_get_timings.argtypes = [vp, vp]
    L.aobake_get_stats.argtypes = [vp, vp]
    L.aobake_num_samples.argtypes = [vp]
    L.aobake_num_samples.restype = sz
    if path is None:
        _LIB = L
    return L


def default_params() -> AoBakeParams:
    p = AoBakeParams()
    load_library().aobake_default_params(C.byref(p))
    return p


class Baker:
    """Resident bake context (AoBake*)."""

    def __init__(self, device: int = 0, instancing_mode: int = INSTANCING_AUTO, collect_stats: bool = False,
                 cg_tolerance: float = 1e-6, cg_max_iterations: int = 20000, trace_kernel: int = 0,
                 refill_below: int = 0, leaf_tris: int = 0, node_test: int = 0, deferred_capacity: int = 0, tri_batch: int = 0,
                 no_oversized_split: bool = False, ls_energy: int = 0, ls_matrix_free: bool = False, ray_order: int = 0):
        self.lib = load_library()
        p = default_params()
        p.device, p.instancing_mode, p.collect_stats = device, instancing_mode, int(collect_stats)
        p.cg_tolerance, p.cg_max_iterations, p.trace_kernel = cg_tolerance, cg_max_iterations, trace_kernel
        p.refill_below = refill_below
        p.leaf_tris = leaf_tris
        p.node_test = node_test
        p.deferred_capacity = deferred_capacity
        p.tri_batch = tri_batch
        p.no_oversized_split = int(no_oversized_split)
        p.ls_energy = ls_energy
        p.ls_matrix_free = int(ls_matrix_free)
        p.ray_order = int(ray_order)
        self.device = device
        self._h = C.c_void_p()
        rc = self.lib.aobake_create(C.byref(p), C.byref(self._h))
        if rc != 0:
            raise AoBakeError(rc, (self.lib.aobake_last_error(None) or b"").decode())
        self.scene: Optional[Scene] = None
        self.per_instance: Optional[np.ndarray] = None
        self._keep = []

    # -- plumbing --
    def _ck(self, rc: int):
        if rc != 0:
            raise AoBakeError(rc, (self.lib.aobake_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.aobake_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self.lib.aobake_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._ck(self.lib.aobake_synchronize(self._h))

    # -- bake path --
    def set_scene(self, scene: Scene, blockers: Optional[Scene] = None, distributed: bool = False):
        """distributed=True: every rank of comm_init holds the same host scene; each uploads 1/N of it and
        NCCL all-gathers the rest (aobake_set_scene_distributed)."""
        ps = PackedScene(scene)
        pb = PackedScene(blockers) if blockers is not None and len(blockers.instances) else None
        fn = self.lib.aobake_set_scene_distributed if distributed else self.lib.aobake_set_scene
        self._ck(fn(self._h, ps.ref(), pb.ref() if pb else None))
        self.scene = scene
        self.per_instance = None

    def set_scene_geometry(self, scene: Scene):
        """Scene upload without a BVH: for sampling and the vertex maps only."""
        ps = PackedScene(scene)
        self._ck(self.lib.aobake_set_scene_geometry(self._h, ps.ref()))
        self.scene = scene
        self.per_instance = None

    def distribute_samples(self, min_samples_per_triangle: int, requested_num_samples: int):
        n = len(self.scene.instances)
        per = (C.c_size_t * max(n, 1))()
        total = C.c_size_t()
        self._ck(self.lib.aobake_distribute_samples(self._h, min_samples_per_triangle, requested_num_samples, per,
                                                    C.byref(total)))
        return int(total.value), np.array(per[:n], dtype=np.uint64)

    def sample_instances(self, per_instance: Sequence[int], min_samples_per_triangle: int,
                         download: bool = True) -> Optional[SampleBuffers]:
        per = (C.c_size_t * max(len(per_instance), 1))(*[int(x) for x in per_instance])
        sb = SampleBuffers(int(sum(int(x) for x in per_instance))) if download else None
        self._ck(self.lib.aobake_sample_instances(self._h, per, min_samples_per_triangle, sb.ref() if sb else None))
        self.per_instance = np.array([int(x) for x in per_instance], dtype=np.uint64)
        return sb

    def set_samples(self, samples: SampleBuffers, per_instance: Optional[Sequence[int]] = None, distributed: bool = False):
        """distributed=True: upload only the super-blocks this rank traces in compute_ao_distributed."""
        per = None
        if per_instance is not None:
            per = (C.c_size_t * max(len(per_instance), 1))(*[int(x) for x in per_instance])
            self.per_instance = np.array([int(x) for x in per_instance], dtype=np.uint64)
        fn = self.lib.aobake_set_samples_distributed if distributed else self.lib.aobake_set_samples
        self._ck(fn(self._h, samples.ref(), per))

    @property
    def num_samples(self) -> int:
        return int(self.lib.aobake_num_samples(self._h))

    def compute_ao(self, rays_per_sample: int, scene_offset: float, scene_maxdistance: float, download: bool = True,
                   begin: Optional[int] = None, end: Optional[int] = None, out: Optional[np.ndarray] = None):
        n = self.num_samples
        b = 0 if begin is None else begin
        e = n if end is None else end
        ao = None
        if download:
            ao = out if out is not None else np.empty(e - b, dtype=np.float32)
            assert ao.dtype == np.float32 and ao.size >= e - b
        self._ck(self.lib.aobake_compute_ao_range(self._h, b, e, rays_per_sample, float(scene_offset),
                                                  float(scene_maxdistance), ao.ctypes.data if ao is not None else None))
        return ao

    def compute_ao_interleaved(self, part: int, num_parts: int, rays_per_sample: int, scene_offset: float,
                               scene_maxdistance: float, block_samples: int = 0):
        """Trace the super-blocks owned by `part` of `num_parts`; other samples' AO is set to 0."""
        self._ck(self.lib.aobake_compute_ao_interleaved(self._h, part, num_parts, block_samples, rays_per_sample,
                                                        float(scene_offset), float(scene_maxdistance)))

    # -- native NCCL exchange (inside libaobake.so) --
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load_library().aobake_comm_unique_id(buf)
        if rc != 0:
            raise AoBakeError(rc, (load_library().aobake_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        assert len(unique_id) == 128
        self._ck(self.lib.aobake_comm_init(self._h, rank, nranks, C.create_string_buffer(unique_id, 128)))

    def comm_destroy(self):
        self._ck(self.lib.aobake_comm_destroy(self._h))

    def compute_ao_distributed(self, rays_per_sample: int, scene_offset: float, scene_maxdistance: float,
                               download: bool = True, out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
        """Interleaved partition over the ranks of comm_init + in-place ncclAllReduce, all native."""
        ao = (out if out is not None else np.empty(self.num_samples, dtype=np.float32)) if download else None
        self._ck(self.lib.aobake_compute_ao_distributed(self._h, rays_per_sample, float(scene_offset), float(scene_maxdistance),
                                                        ao.ctypes.data if ao is not None else None))
        return ao

    def download_ao(self) -> np.ndarray:
        import ctypes as _C
        out = np.empty(self.num_samples, dtype=np.float32)
        ptr, n = self.ao_device_ptr()
        if n:
            from .multi_gpu import device_tensor
            out[...] = device_tensor(ptr, n, self.device).cpu().numpy()
        return out

    def ao_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.aobake_get_ao_device(self._h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def set_ao(self, ao: np.ndarray):
        ao = np.ascontiguousarray(ao, dtype=np.float32)
        assert ao.size == self.num_samples
        self._ck(self.lib.aobake_set_ao(self._h, ao.ctypes.data))

    def hit_counts(self) -> np.ndarray:
        out = np.zeros(self.num_samples, dtype=np.uint32)
        self._ck(self.lib.aobake_get_hit_counts(self._h, out.ctypes.data))
        return out

    def map_ao_to_vertices(self, mode: int = FILTER_AREA_BASED, regularization_weight: float = 0.1,
                           distributed: bool = False) -> List[np.ndarray]:
        """distributed=True: split the instances over the ranks of comm_init (native NCCL gather)."""
        arrs = [np.zeros(len(self.scene.meshes[i.mesh_index].vertices), dtype=np.float32) for i in self.scene.instances]
        ptrs = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
        fn = self.lib.aobake_map_ao_to_vertices_distributed if distributed else self.lib.aobake_map_ao_to_vertices
        self._ck(fn(self._h, mode, float(regularization_weight), ptrs))
        return arrs

    # -- parity hooks --
    def trace_rays(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hit = np.zeros(len(rays), dtype=np.uint8)
        self._ck(self.lib.aobake_trace_rays(self._h, rays.ctypes.data, len(rays), hit.ctypes.data))
        return hit

    def dump_rays(self, begin: int, end: int, rays_per_sample: int, scene_offset: float, scene_maxdistance: float):
        q = int(np.float32(np.sqrt(np.float32(rays_per_sample))) + np.float32(0.5))
        out = np.zeros((end - begin, q * q, 8), dtype=np.float32)
        self._ck(self.lib.aobake_dump_rays(self._h, begin, end, rays_per_sample, float(scene_offset),
                                           float(scene_maxdistance), out.ctypes.data))
        return out

    def timings(self) -> AoTimings:
        t = AoTimings()
        self._ck(self.lib.aobake_get_timings(self._h, C.byref(t)))
        return t

    def stats(self) -> AoStats:
        s = AoStats()
        self._ck(self.lib.aobake_get_stats(self._h, C.byref(s)))
        return s


# ---- one-shot forms with the reference's names and argument order (bake_api.h) ------------
def distributeSamples(scene: Scene, min_samples_per_triangle: int, requested_num_samples: int, device: int = 0):
    with Baker(device) as b:
        b.set_scene_geometry(scene)
        return b.distribute_samples(min_samples_per_triangle, requested_num_samples)


def sampleInstances(scene: Scene, num_samples_per_instance, min_samples_per_triangle: int, device: int = 0) -> SampleBuffers:
    with Baker(device) as b:
        b.set_scene_geometry(scene)
        return b.sample_instances(num_samples_per_instance, min_samples_per_triangle)


def computeAO(scene: Scene, blockers: Optional[Scene], ao_samples: SampleBuffers, rays_per_sample: int,
              scene_offset: float, scene_maxdistance: float, device: int = 0) -> np.ndarray:
    """Like the reference, builds the acceleration structure inside the call."""
    with Baker(device) as b:
        b.set_scene(scene, blockers)
        b.set_samples(ao_samples)
        return b.compute_ao(rays_per_sample, scene_offset, scene_maxdistance)


def mapAOToVertices(scene: Scene, num_samples_per_instance, ao_samples: SampleBuffers, ao_values: np.ndarray,
                    mode: int = FILTER_LEAST_SQUARES, regularization_weight: float = 0.1, device: int = 0):
    with Baker(device) as b:
        b.set_scene_geometry(scene)
        b.set_samples(ao_samples, num_samples_per_instance)
        b.set_ao(ao_values)
        return b.map_ao_to_vertices(mode, regularization_weight)


def make_ground_plane(bbox_min, bbox_max, upaxis=1, scale_factor=100.0, offset_factor=0.03):
    lo = np.ascontiguousarray(bbox_min, dtype=np.float32)
    hi = np.ascontiguousarray(bbox_max, dtype=np.float32)
    v = np.zeros((4, 3), dtype=np.float32)
    t = np.zeros((2, 3), dtype=np.uint32)
    rc = load_library().aobake_make_ground_plane(lo.ctypes.data, hi.ctypes.data, upaxis, scale_factor, offset_factor,
                                                 v.ctypes.data, t.ctypes.data)
    if rc != 0:
        raise AoBakeError(rc, "make_ground_plane")
    return v, t
